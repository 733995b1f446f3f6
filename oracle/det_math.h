// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under royaltracer-dx_b200/ may include this file.
//
// det_math.h: the numerics contract of the CPU oracle.
//
// The reference's HLSL (DXIL, not `precise`) leaves sin/cos/rsqrt/pow loosely specified
// (SURVEY.md Appendix C.3), so its own GPU output is not bit-defined.  The oracle therefore
// *defines* every transcendental as a fixed sequence of IEEE-754 binary32 +,-,*,/ and sqrt
// operations evaluated left to right with no FMA contraction (compile with
// -ffp-contract=off, no -ffast-math).  The CUDA product restates the same sequences
// (csrc/dmath.cuh, compiled with -fmad=false) and must agree bit for bit.
//
// HLSL intrinsic semantics followed here:
//   saturate(NaN)=0, min/max(NaN,x)=x (fminf/fmaxf), normalize(v)=v*rsqrt(dot(v,v)),
//   rsqrt(x)=1/sqrt(x), pow(x,5)=x*x*x*x*x, half = IEEE binary16 with RNE conversions.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

struct f3 { float x, y, z; };
struct f4 { float x, y, z, w; };

static inline f3 mk3(float x, float y, float z) { f3 r = {x, y, z}; return r; }
static inline f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
static inline f3 operator*(float s, f3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
static inline f3 operator/(f3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
static inline f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }

static inline float dot3(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline f3 cross3(f3 a, f3 b) {
    return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float d_rsqrt(float x) { return 1.0f / sqrtf(x); }
static inline float length3(f3 a) { return sqrtf(dot3(a, a)); }
static inline f3 normalize3(f3 a) { return a * d_rsqrt(dot3(a, a)); }
static inline float saturate1(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
static inline f3 saturate3(f3 a) { return mk3(saturate1(a.x), saturate1(a.y), saturate1(a.z)); }
static inline float lerp1(float a, float b, float t) { return a + t * (b - a); }
static inline f3 reflect3(f3 i, f3 n) { return i - (2.0f * dot3(n, i)) * n; }
static inline bool isnan1(float x) { return x != x; }
static inline bool isinf1(float x) { return fabsf(x) == INFINITY; }
static inline bool any_nan_inf(f3 a) {
    return isnan1(a.x) || isnan1(a.y) || isnan1(a.z) || isinf1(a.x) || isinf1(a.y) || isinf1(a.z);
}

// HLSL mul(M, v) for a float4x4 whose 64 bytes were written by the host as an XMMATRIX /
// glm matrix and are read by HLSL with the default column-major packing:
//   M[r][c] = mem[4*c + r]   
static inline f4 mul44(const float* m, float x, float y, float z, float w) {
    f4 r;
    r.x = ((m[0] * x + m[4] * y) + m[8] * z) + m[12] * w;
    r.y = ((m[1] * x + m[5] * y) + m[9] * z) + m[13] * w;
    r.z = ((m[2] * x + m[6] * y) + m[10] * z) + m[14] * w;
    r.w = ((m[3] * x + m[7] * y) + m[11] * z) + m[15] * w;
    return r;
}

// ---------------------------------------------------------------- binary16 (half)
// Round-to-nearest-even float -> half -> float, by bit manipulation (no F16C dependence).
static inline uint16_t f2h_bits(float f) {
    uint32_t x; memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t ax = x & 0x7fffffffu;
    if (ax >= 0x7f800000u) {                      // inf / nan
        return (uint16_t)(sign | 0x7c00u | ((ax > 0x7f800000u) ? 0x200u : 0u));
    }
    if (ax >= 0x477ff000u) {                      // >= 65520 rounds to inf
        return (uint16_t)(sign | 0x7c00u);
    }
    if (ax < 0x38800000u) {                       // below the smallest normal half (2^-14)
        if (ax < 0x33000000u) return (uint16_t)sign;   // < 2^-25 -> 0 (2^-25 itself ties to even = 0)
        // subnormal: value = mant * 2^(e-150); half subnormal quantum 2^-24
        uint32_t e = ax >> 23;
        uint32_t mant = (ax & 0x7fffffu) | 0x800000u;
        uint32_t shift = 126u - e;                // 14..24 (number of bits dropped from the 24-bit mant)
        uint32_t q = mant >> shift;
        uint32_t rem = mant & ((1u << shift) - 1u);
        uint32_t half = 1u << (shift - 1u);
        if (rem > half || (rem == half && (q & 1u))) q++;
        return (uint16_t)(sign | q);
    }
    uint32_t e = (ax >> 23) - 112u;               // rebias 127 -> 15
    uint32_t mant = ax & 0x7fffffu;
    uint32_t q = (e << 10) | (mant >> 13);
    uint32_t rem = mant & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (q & 1u))) q++;   // carry may bump the exponent: correct
    return (uint16_t)(sign | q);
}
static inline float h2f_bits(uint16_t h) {
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    uint32_t e = (h >> 10) & 0x1fu;
    uint32_t m = h & 0x3ffu;
    uint32_t x;
    if (e == 0) {
        if (m == 0) x = sign;
        else {
            float v = (float)m * 5.9604644775390625e-8f;   // m * 2^-24, exact
            memcpy(&x, &v, 4); x |= sign;
        }
    } else if (e == 31) x = sign | 0x7f800000u | (m << 13);
    else x = sign | ((e + 112u) << 23) | (m << 13);
    float f; memcpy(&f, &x, 4); return f;
}
// q16(x): the float value of half(x)
static inline float q16(float x) { return h2f_bits(f2h_bits(x)); }
static inline f3 q16v(f3 a) { return mk3(q16(a.x), q16(a.y), q16(a.z)); }
// half*half -> half: the product of two binary16 values is exact in binary32.
static inline float hmul(float a, float b) { return q16(a * b); }
// half+half -> half: the sum of two binary16 values is exact in binary64; round once.
static inline float hadd(float a, float b) {
    double s = (double)a + (double)b;
    float f = (float)s;
    // (float)s is exact unless the sum needs > 24 bits; in that case re-round correctly:
    if ((double)f != s) {
        // s lies strictly between two floats; pick by sticky bit: nudge f toward s so that the
        // subsequent RNE to 11 bits sees the right side of any tie.
        uint32_t u; memcpy(&u, &f, 4);
        // make the float odd in its last bit (round-to-odd) — a standard double-rounding cure
        if ((u & 1u) == 0u) {
            bool up = (s > (double)f);
            if (f >= 0.0f) u = up ? u + 1u : u - 1u; else u = up ? u - 1u : u + 1u;
            memcpy(&f, &u, 4);
        }
    }
    return q16(f);
}

// ---------------------------------------------------------------- sin/cos
// Cephes-style single precision sincos, restricted to what the reference needs
// (arguments 2*PI*u, u in [0,1]); defined for |x| < 8192.
static inline void d_sincos(float x, float* s_out, float* c_out) {
    float ax = fabsf(x);
    int j = (int)(ax * 1.27323954473516f);       // 4/pi
    j = (j + 1) & ~1;                            // nearest even octant boundary
    float y = (float)j;
    float z = ((ax - y * 0.78515625f) - y * 2.4187564849853515625e-4f) - y * 3.77489497744594108e-8f;
    float zz = z * z;
    float sp = ((-1.9515295891e-4f * zz + 8.3321608736e-3f) * zz - 1.6666654611e-1f) * zz * z + z;
    float cp = ((2.443315711809948e-5f * zz - 1.388731625493765e-3f) * zz + 4.166664568298827e-2f) * zz * zz
               - 0.5f * zz + 1.0f;
    int q = (j >> 1) & 3;                        // quadrant
    float s, c;
    switch (q) {
        case 0: s = sp;  c = cp;  break;
        case 1: s = cp;  c = -sp; break;
        case 2: s = -sp; c = -cp; break;
        default: s = -cp; c = sp; break;
    }
    if (x < 0.0f) s = -s;
    *s_out = s; *c_out = c;
}

// ---------------------------------------------------------------- log/exp/pow (sRGB only)
static inline float d_log(float x) {             // x > 0, finite, normal
    uint32_t u; memcpy(&u, &x, 4);
    int e = (int)((u >> 23) & 0xffu) - 126;      // x = m * 2^e, m in [0.5,1)
    u = (u & 0x007fffffu) | 0x3f000000u;
    float m; memcpy(&m, &u, 4);
    if (m < 0.707106781186547524f) { e -= 1; m = (m + m) - 1.0f; } else { m = m - 1.0f; }
    float z = m * m;
    float y = ((((((((7.0376836292e-2f * m - 1.1514610310e-1f) * m + 1.1676998740e-1f) * m
                 - 1.2420140846e-1f) * m + 1.4249322787e-1f) * m - 1.6668057665e-1f) * m
                 + 2.0000714765e-1f) * m - 2.4999993993e-1f) * m + 3.3333331174e-1f) * m * z;
    float fe = (float)e;
    y = y + -2.12194440e-4f * fe;
    y = y + -0.5f * z;
    float r = m + y;
    r = r + 0.693359375f * fe;
    return r;
}
static inline float d_exp(float x) {             // |x| < 80
    float fz = floorf(1.44269504088896341f * x + 0.5f);
    int n = (int)fz;
    x = x - fz * 0.693359375f;
    x = x - fz * -2.12194440e-4f;
    float z = x * x;
    float p = (((((1.9875691500e-4f * x + 1.3981999507e-3f) * x + 8.3334519073e-3f) * x
                + 4.1665795894e-2f) * x + 1.6666665459e-1f) * x + 5.0000001201e-1f) * z + x + 1.0f;
    uint32_t sc = (uint32_t)(n + 127) << 23;     // 2^n, n in the normal range
    float s; memcpy(&s, &sc, 4);
    return p * s;
}
static inline float d_pow(float x, float y) { return d_exp(d_log(x) * y); }

}  // namespace orc
