/* Exhaustive check behind DESIGN.md section 8 (K2): can an integer expression replace OpenCV's float32 blend in the four
 * diagonal LBP bits?  For diagonal n the reference computes, in float32, left to right and without FMA,
 *     t = ((w1*p1 + w2*p2) + w3*p3) + w4*p4,   bit = (t > c) || (fabsf(t - c) < FLT_EPSILON)
 * where one of p1..p4 is the centre pixel c itself.  In exact arithmetic
 *     t - c = A*(x + y) + B*z - (1 - C)*c,   A = 0x3e5413cd, B = 0x3effffff, C = 0x3dafb0ce
 * (x, y: the two taps of weight A, z: the tap of weight B), i.e. h = 434334*(x + y) + 1048576*z - 1917244*c scaled by 2^21
 * with the weights rounded to integers.  This program walks all 2^32 (x, y, z, c) for each of the four summation orders
 * and reports the smallest |h| at which sign(h) and the float decision DISAGREE ... and the largest |h| at which they
 * disagree: below that magnitude the float rounding (and the FLT_EPSILON clause) decides, above it h alone does.
 * Build: gcc -O2 -fopenmp -ffp-contract=off -o /tmp/lbp_integer_form profiles/micro/lbp_integer_form.c ; run: ~1 min. */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static float f32(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

int main(void)
{
    const float A = f32(0x3e5413cdu), B = f32(0x3effffffu), C = f32(0x3dafb0ceu);
    /* order[n]: which of (x, y, z, c) sits at position 1..4 and with which weight; 0=x(A) 1=y(A) 2=z(B) 3=c(C) */
    static const int order[4][4] = {
        {0, 2, 3, 1},   /* n = 1:  A*N  + B*NE + C*c  + A*E  */
        {2, 0, 1, 3},   /* n = 3:  B*NW + A*N  + A*W  + C*c  */
        {0, 3, 2, 1},   /* n = 5:  A*W  + C*c  + B*SW + A*S  */
        {3, 0, 1, 2},   /* n = 7:  C*c  + A*E  + A*S  + B*SE */
    };
    for (int n = 0; n < 4; n++) {
        long long worst = -1;          /* largest |h| with a disagreement */
        long long ties = 0, disagreements = 0;
#pragma omp parallel for reduction(max : worst) reduction(+ : ties, disagreements) schedule(static)
        for (int c = 0; c < 256; c++)
            for (int x = 0; x < 256; x++)
                for (int y = 0; y < 256; y++)
                    for (int z = 0; z < 256; z++) {
                        const float p[4] = {(float)x, (float)y, (float)z, (float)c};
                        const float w[4] = {A, A, B, C};
                        float t = w[order[n][0]] * p[order[n][0]];
                        t = t + w[order[n][1]] * p[order[n][1]];
                        t = t + w[order[n][2]] * p[order[n][2]];
                        t = t + w[order[n][3]] * p[order[n][3]];
                        const int bit = (t > (float)c) || (fabsf(t - (float)c) < FLT_EPSILON);
                        const long long h = 434334LL * (x + y) + 1048576LL * z - 1917244LL * c;
                        if (fabsf(t - (float)c) < FLT_EPSILON) ties++;
                        if ((h >= 0) != bit) {
                            disagreements++;
                            const long long a = h < 0 ? -h : h;
                            if (a > worst) worst = a;
                        }
                    }
        printf("diagonal n=%d: %lld float near-ties (|t-c| < FLT_EPSILON), %lld inputs where sign(h) != float bit, largest |h| among them %lld\n",
               2 * n + 1, ties, disagreements, worst);
    }
    return 0;
}
