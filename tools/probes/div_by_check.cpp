// Host check of vote_center.cu::div_by (reciprocal + two FMA corrections) against IEEE division, 480 M numerators over 8 voxel sizes:
//   g++ -O2 -ffp-contract=off -o div_by_check tools/probes/div_by_check.cpp -lm && ./div_by_check      -> 0 mismatches
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
// host emulation of div_by with fmaf (correctly rounded FMA in glibc) against IEEE division
static float div_by(float a, float b, float inv_b) { float q = a * inv_b; float r = fmaf(-b, q, a); q = fmaf(r, inv_b, q); r = fmaf(-b, q, a); return fmaf(r, inv_b, q); }
int main() {
    const float divisors[] = {0.002f, 0.01f, 0.0015f, 0.004f, 0.003f, 0.0025f, 0.00123f, 0.05f};
    long long bad = 0, n = 0;
    uint64_t s = 88172645463325252ull;
    for (float b : divisors) {
        const float inv = 1.0f / b;     // correctly rounded
        for (long long i = 0; i < 60000000; ++i) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            // numerators like (c + offset) - lo: 0 .. 1.2 m, both signs, plus values packed near cell boundaries
            float a = (float)((s >> 11) * (1.0 / 9007199254740992.0)) * 1.2f;
            if (i & 1) a = -a * 0.01f;
            if ((i & 7) == 3) a = b * (float)((s >> 40) % 600) + b * 0.5f + (float)((int)(s & 255) - 128) * 1e-9f;
            const float want = a / b, got = div_by(a, b, inv);
            if (!(want == got)) { if (bad < 5) printf("b=%g a=%.9g want=%.9g got=%.9g\n", b, a, want, got); ++bad; }
            ++n;
        }
    }
    printf("%lld mismatches of %lld\n", bad, n);
    return bad != 0;
}
