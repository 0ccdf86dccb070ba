#!/bin/bash
# static SASS opcode mix of one kernel: tools/sass_mix.sh <object> <mangled-name-regex>
obj=$1; pat=$2
cuobjdump -sass "$obj" | awk -v pat="$pat" '/Function :/{f=($0 ~ pat)} f' > /tmp/_k.sass
n=$(grep -cE "^\s+/\*[0-9a-f]{4}\*/" /tmp/_k.sass)
echo "instructions: $n"
grep -E "^\s+/\*[0-9a-f]{4}\*/" /tmp/_k.sass | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//' | sed -E 's/^@!?U?P[0-9T]+ //' | awk '{print $1}' | sed -E 's/^([A-Z0-9_]+)(\.(128|64|E|F32|RD|RZ|SAT|POPC))?.*/\1\2/' | sort | uniq -c | sort -rn | head -${3:-30}
