#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <omp.h>
int main() {
    const float r = 1.0f / 255.0f;                       // RN(1/255)
    const float r2 = (float)(1.0 / 255.0 - (double)r);  // RN of the residual
    printf("r=%.10g r2=%.10g\n", r, r2);
    uint32_t hi; float top = 256.0f; memcpy(&hi, &top, 4);
    long bad2 = 0, bad3 = 0;
#pragma omp parallel for reduction(+:bad2,bad3)
    for (uint32_t u = 0; u <= hi; ++u) {
        float x; memcpy(&x, &u, 4);
        float ref = x / 255.0f;
        float q2 = fmaf(x, r, x * r2);
        float q0 = x * r;
        float q3 = fmaf(fmaf(-255.0f, q0, x), r, q0);
        if (q2 != ref) bad2++;
        if (q3 != ref) bad3++;
    }
    printf("bad 2-op: %ld  bad 3-op (Markstein): %ld of %u\n", bad2, bad3, hi + 1);
    return 0;
}
