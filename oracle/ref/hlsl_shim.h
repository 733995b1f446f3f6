// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under royaltracer-dx_b200/ may include this file.
//
// hlsl_shim.h: just enough of HLSL (SM 6.7, -enable-16bit-types, DXR 1.0 intrinsics) in C++17 for g++ to compile the reference's
// shader files (/root/reference/Pathtracer/{shaders,include}/*.hlsl) after the mechanical token filter of make_ref.py, so that the
// reference's OWN TEXT can be executed on the CPU and the hand-written oracle (oracle/rtx_oracle.cpp) can be pinned against it.
//
// What this header defines is the part HLSL leaves to the compiler / the hardware — the numerics contract of oracle/det_math.h:
//   * every float operation is IEEE binary32, evaluated in source order, no FMA contraction (-ffp-contract=off);
//   * dot(a,b) = (a.x*b.x + a.y*b.y) + a.z*b.z (+ a.w*b.w), normalize(v) = v * rsqrt(dot(v,v)), rsqrt(x) = 1/sqrt(x),
//     length(v) = sqrt(dot(v,v)), reflect(i,n) = i - (2*dot(n,i))*n, lerp(a,b,t) = a + t*(b-a), saturate(NaN) = 0,
//     min/max(NaN,x) = x, pow(x,5) = x*x*x*x*x, pow(x,1) = x, other pow = exp(log(x)*y) (det_math), sin/cos = det_math d_sincos;
//   * mul(M, v) with M read column-major from the 64 bytes the host wrote: M[r][c] = mem[4c + r], summed left to right;
//   * half = IEEE binary16 with round-to-nearest-even conversions; half (op) half is rounded once to binary16;
//   * an out-of-bounds StructuredBuffer read returns zeros (D3D12 robust buffer access), an uninitialised local is zero.
// Everything else — control flow, formulae, constants, the order of the random draws — comes from the reference's files.
#pragma once
#include <stdint.h>
#include <string.h>

#include <type_traits>

#include "../det_math.h"

namespace hlsl {

typedef unsigned int uint;
typedef unsigned short uint16_t;

// ------------------------------------------------------------------------------------------------ half
struct half {
    float v;                                    // always a binary16-representable value
    half() = default;
    half(float f) : v(orc::q16(f)) {}
    half(double f) : v(orc::q16((float)f)) {}
    half(int i) : v(orc::q16((float)i)) {}
    half(uint i) : v(orc::q16((float)i)) {}
    operator float() const { return v; }
};
inline half h_raw(float exact) { half h; h.v = exact; return h; }
inline half operator+(half a, half b) { return h_raw(orc::hadd(a.v, b.v)); }
inline half operator-(half a, half b) { return h_raw(orc::hadd(a.v, -b.v)); }
inline half operator*(half a, half b) { return h_raw(orc::hmul(a.v, b.v)); }
inline half operator/(half a, half b) { return half(a.v / b.v); }
inline half operator-(half a) { return h_raw(-a.v); }
inline bool operator<(half a, half b) { return a.v < b.v; }
inline bool operator>(half a, half b) { return a.v > b.v; }
inline bool operator<=(half a, half b) { return a.v <= b.v; }
inline bool operator>=(half a, half b) { return a.v >= b.v; }
inline bool operator==(half a, half b) { return a.v == b.v; }
inline bool operator!=(half a, half b) { return a.v != b.v; }

template <class S> struct is_scalar : std::is_arithmetic<S> {};
template <> struct is_scalar<half> : std::true_type {};
// usual arithmetic conversions, extended by half (half with an integer stays half, with float/double promotes)
template <class A, class B> struct promote { typedef typename std::common_type<A, B>::type type; };
template <> struct promote<half, half> { typedef half type; };
template <class B> struct promote<half, B> { typedef typename std::conditional<std::is_floating_point<B>::value, B, half>::type type; };
template <class A> struct promote<A, half> { typedef typename std::conditional<std::is_floating_point<A>::value, A, half>::type type; };
#define HLSL_HALF_MIXED(OP)                                                                                                   \
    template <class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type>                                   \
    inline typename promote<half, S>::type operator OP(half a, S b) { typedef typename promote<half, S>::type R; return R(a) OP R(b); } \
    template <class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type>                                   \
    inline typename promote<S, half>::type operator OP(S a, half b) { typedef typename promote<S, half>::type R; return R(a) OP R(b); }
HLSL_HALF_MIXED(+) HLSL_HALF_MIXED(-) HLSL_HALF_MIXED(*) HLSL_HALF_MIXED(/)
#undef HLSL_HALF_MIXED
#define HLSL_HALF_CMP(OP)                                                                                                     \
    template <class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type>                                   \
    inline bool operator OP(half a, S b) { return a.v OP (float)b; }                                                          \
    template <class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type>                                   \
    inline bool operator OP(S a, half b) { return (float)a OP b.v; }
HLSL_HALF_CMP(<) HLSL_HALF_CMP(>) HLSL_HALF_CMP(<=) HLSL_HALF_CMP(>=) HLSL_HALF_CMP(==) HLSL_HALF_CMP(!=)
#undef HLSL_HALF_CMP

template <class To, class From> inline To conv(From f) { return (To)f; }
template <> inline bool conv<bool, half>(half f) { return f.v != 0.0f; }

// ------------------------------------------------------------------------------------------------ vectors
template <class T, int N> struct vec;

template <class T> struct vec<T, 2> {
    union { T x; T r; }; union { T y; T g; };
    vec() : x(T(0)), y(T(0)) {}
    vec(T s) : x(s), y(s) {}
    template <class A, class B, class = typename std::enable_if<is_scalar<A>::value && is_scalar<B>::value>::type>
    vec(A a, B b) : x(conv<T>(a)), y(conv<T>(b)) {}
    template <class U> vec(const vec<U, 2>& o) : x(conv<T>(o.x)), y(conv<T>(o.y)) {}
    template <class U> vec(const vec<U, 3>& o) : x(conv<T>(o.x)), y(conv<T>(o.y)) {}       // HLSL implicit truncation
    T& operator[](int i) { return i == 0 ? x : y; }
    T operator[](int i) const { return i == 0 ? x : y; }
    vec<T, 2> xy() const { return *this; }
    operator T() const { return x; }            // HLSL implicit truncation to a scalar (warning X3206)
};
template <class T> struct vec<T, 3> {
    union { T x; T r; }; union { T y; T g; }; union { T z; T b; };
    vec() : x(T(0)), y(T(0)), z(T(0)) {}
    vec(T s) : x(s), y(s), z(s) {}
    template <class A, class B, class C, class = typename std::enable_if<is_scalar<A>::value && is_scalar<B>::value && is_scalar<C>::value>::type>
    vec(A a, B b_, C c) : x(conv<T>(a)), y(conv<T>(b_)), z(conv<T>(c)) {}
    template <class U, class C, class = typename std::enable_if<is_scalar<C>::value>::type>
    vec(const vec<U, 2>& a, C c) : x(conv<T>(a.x)), y(conv<T>(a.y)), z(conv<T>(c)) {}
    template <class U> vec(const vec<U, 3>& o) : x(conv<T>(o.x)), y(conv<T>(o.y)), z(conv<T>(o.z)) {}
    template <class U> vec(const vec<U, 4>& o) : x(conv<T>(o.x)), y(conv<T>(o.y)), z(conv<T>(o.z)) {}   // implicit truncation
    T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    T operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    vec<T, 2> xy() const { return vec<T, 2>(x, y); }
    vec<T, 3> xyz() const { return *this; }
    operator T() const { return x; }
};
template <class T> struct vec<T, 4> {
    union { T x; T r; }; union { T y; T g; }; union { T z; T b; }; union { T w; T a; };
    vec() : x(T(0)), y(T(0)), z(T(0)), w(T(0)) {}
    vec(T s) : x(s), y(s), z(s), w(s) {}
    template <class A, class B, class C, class D,
              class = typename std::enable_if<is_scalar<A>::value && is_scalar<B>::value && is_scalar<C>::value && is_scalar<D>::value>::type>
    vec(A a_, B b_, C c, D d) : x(conv<T>(a_)), y(conv<T>(b_)), z(conv<T>(c)), w(conv<T>(d)) {}
    template <class U, class D, class = typename std::enable_if<is_scalar<D>::value>::type>
    vec(const vec<U, 3>& v, D d) : x(conv<T>(v.x)), y(conv<T>(v.y)), z(conv<T>(v.z)), w(conv<T>(d)) {}
    template <class U, class C, class D, class = typename std::enable_if<is_scalar<C>::value && is_scalar<D>::value>::type>
    vec(const vec<U, 2>& v, C c, D d) : x(conv<T>(v.x)), y(conv<T>(v.y)), z(conv<T>(c)), w(conv<T>(d)) {}
    template <class U> vec(const vec<U, 4>& o) : x(conv<T>(o.x)), y(conv<T>(o.y)), z(conv<T>(o.z)), w(conv<T>(o.w)) {}
    T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    T operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    vec<T, 2> xy() const { return vec<T, 2>(x, y); }
    vec<T, 3> xyz() const { return vec<T, 3>(x, y, z); }
    void set_xyz(const vec<T, 3>& v) { x = v.x; y = v.y; z = v.z; }
    operator T() const { return x; }
};

typedef vec<float, 2> float2; typedef vec<float, 3> float3; typedef vec<float, 4> float4;
typedef vec<half, 2> half2; typedef vec<half, 3> half3; typedef vec<half, 4> half4;
typedef vec<uint, 2> uint2; typedef vec<uint, 3> uint3; typedef vec<uint, 4> uint4;
typedef vec<int, 2> int2; typedef vec<int, 3> int3; typedef vec<int, 4> int4;
typedef vec<bool, 2> bool2; typedef vec<bool, 3> bool3; typedef vec<bool, 4> bool4;

// component-wise map helpers
template <class R, class T, class F> inline vec<R, 2> map1(const vec<T, 2>& a, F f) { return vec<R, 2>(f(a.x), f(a.y)); }
template <class R, class T, class F> inline vec<R, 3> map1(const vec<T, 3>& a, F f) { return vec<R, 3>(f(a.x), f(a.y), f(a.z)); }
template <class R, class T, class F> inline vec<R, 4> map1(const vec<T, 4>& a, F f) { return vec<R, 4>(f(a.x), f(a.y), f(a.z), f(a.w)); }
template <class R, class T, class U, class F> inline vec<R, 2> map2(const vec<T, 2>& a, const vec<U, 2>& b, F f) { return vec<R, 2>(f(a.x, b.x), f(a.y, b.y)); }
template <class R, class T, class U, class F> inline vec<R, 3> map2(const vec<T, 3>& a, const vec<U, 3>& b, F f) { return vec<R, 3>(f(a.x, b.x), f(a.y, b.y), f(a.z, b.z)); }
template <class R, class T, class U, class F> inline vec<R, 4> map2(const vec<T, 4>& a, const vec<U, 4>& b, F f) {
    return vec<R, 4>(f(a.x, b.x), f(a.y, b.y), f(a.z, b.z), f(a.w, b.w));
}

#define HLSL_VEC_ARITH(OP)                                                                                                    \
    template <class T, class U, int N> inline vec<typename promote<T, U>::type, N> operator OP(const vec<T, N>& a, const vec<U, N>& b) { \
        typedef typename promote<T, U>::type R;                                                                               \
        return map2<R>(a, b, [](T p, U q) { return R(p) OP R(q); });                                                          \
    }                                                                                                                         \
    template <class T, class S, int N, class = typename std::enable_if<is_scalar<S>::value>::type>                            \
    inline vec<typename promote<T, S>::type, N> operator OP(const vec<T, N>& a, S s) {                                        \
        typedef typename promote<T, S>::type R;                                                                               \
        return map1<R>(a, [s](T p) { return R(p) OP R(s); });                                                                 \
    }                                                                                                                         \
    template <class T, class S, int N, class = typename std::enable_if<is_scalar<S>::value>::type>                            \
    inline vec<typename promote<S, T>::type, N> operator OP(S s, const vec<T, N>& a) {                                        \
        typedef typename promote<S, T>::type R;                                                                               \
        return map1<R>(a, [s](T p) { return R(s) OP R(p); });                                                                 \
    }                                                                                                                         \
    template <class T, class U, int N> inline vec<T, N>& operator OP##=(vec<T, N>& a, const vec<U, N>& b) { a = vec<T, N>(a OP b); return a; } \
    template <class T, class S, int N, class = typename std::enable_if<is_scalar<S>::value>::type>                            \
    inline vec<T, N>& operator OP##=(vec<T, N>& a, S s) { a = vec<T, N>(a OP s); return a; }
HLSL_VEC_ARITH(+) HLSL_VEC_ARITH(-) HLSL_VEC_ARITH(*) HLSL_VEC_ARITH(/)
#undef HLSL_VEC_ARITH
template <class T, int N> inline vec<T, N> operator-(const vec<T, N>& a) { return map1<T>(a, [](T p) { return -p; }); }

#define HLSL_VEC_CMP(OP)                                                                                                      \
    template <class T, class U, int N> inline vec<bool, N> operator OP(const vec<T, N>& a, const vec<U, N>& b) {              \
        return map2<bool>(a, b, [](T p, U q) { return p OP q; });                                                             \
    }                                                                                                                         \
    template <class T, class S, int N, class = typename std::enable_if<is_scalar<S>::value>::type>                            \
    inline vec<bool, N> operator OP(const vec<T, N>& a, S s) { return map1<bool>(a, [s](T p) { return p OP s; }); }
HLSL_VEC_CMP(<) HLSL_VEC_CMP(>) HLSL_VEC_CMP(<=) HLSL_VEC_CMP(>=) HLSL_VEC_CMP(==) HLSL_VEC_CMP(!=)
#undef HLSL_VEC_CMP

inline bool any(bool b) { return b; }
inline bool all(bool b) { return b; }
inline bool any(const bool2& v) { return v.x || v.y; }
inline bool any(const bool3& v) { return v.x || v.y || v.z; }
inline bool any(const bool4& v) { return v.x || v.y || v.z || v.w; }
inline bool all(const bool2& v) { return v.x && v.y; }
inline bool all(const bool3& v) { return v.x && v.y && v.z; }
inline bool all(const bool4& v) { return v.x && v.y && v.z && v.w; }

// ------------------------------------------------------------------------------------------------ scalar intrinsics
inline float sqrt(float x) { return ::sqrtf(x); }
inline float rsqrt(float x) { return orc::d_rsqrt(x); }
inline float abs(float x) { return ::fabsf(x); }
inline int abs(int x) { return x < 0 ? -x : x; }
inline float floor(float x) { return ::floorf(x); }
inline float round(float x) { return ::rintf(x); }              // HLSL round: to nearest even
inline float saturate(float x) { return orc::saturate1(x); }
inline float lerp(float a, float b, float t) { return orc::lerp1(a, b, t); }
inline bool isnan(float x) { return x != x; }
inline bool isinf(float x) { return ::fabsf(x) == INFINITY; }
inline bool isfinite(float x) { return !isnan(x) && !isinf(x); }
inline float sin(float x) { float s, c; orc::d_sincos(x, &s, &c); return s; }
inline float cos(float x) { float s, c; orc::d_sincos(x, &s, &c); return c; }
inline float radians(float d) { return d * 0.017453292f; }
inline float pow(float x, float y) {
    if (y == 5.0f) return x * x * x * x * x;
    if (y == 1.0f) return x;
    return orc::d_pow(x, y);
}
// min / max: NaN-ignoring for floats (fminf/fmaxf), mixed int/float arguments promote
inline float min(float a, float b) { return ::fminf(a, b); }
inline float max(float a, float b) { return ::fmaxf(a, b); }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline uint min(uint a, uint b) { return a < b ? a : b; }
inline uint max(uint a, uint b) { return a > b ? a : b; }
inline float min(int a, float b) { return ::fminf((float)a, b); }
inline float max(int a, float b) { return ::fmaxf((float)a, b); }
inline float min(float a, int b) { return ::fminf(a, (float)b); }
inline float max(float a, int b) { return ::fmaxf(a, (float)b); }
inline float min(double a, float b) { return ::fminf((float)a, b); }
inline float max(double a, float b) { return ::fmaxf((float)a, b); }
inline float min(float a, double b) { return ::fminf(a, (float)b); }
inline float max(float a, double b) { return ::fmaxf(a, (float)b); }
inline int min(int a, uint b) { return min(a, (int)b); }
inline int min(uint a, int b) { return min((int)a, b); }
inline int max(int a, uint b) { return max(a, (int)b); }
inline int max(uint a, int b) { return max((int)a, b); }
inline float clamp(float x, float a, float b) { return min(max(x, a), b); }
inline int clamp(int x, int a, int b) { return min(max(x, a), b); }

// ------------------------------------------------------------------------------------------------ vector intrinsics
#define HLSL_MAP_F(NAME)                                                                                                      \
    template <int N> inline vec<float, N> NAME(const vec<float, N>& a) { return map1<float>(a, [](float p) { return NAME(p); }); }
HLSL_MAP_F(sqrt) HLSL_MAP_F(abs) HLSL_MAP_F(floor) HLSL_MAP_F(round) HLSL_MAP_F(saturate)
#undef HLSL_MAP_F
template <int N> inline vec<bool, N> isnan(const vec<float, N>& a) { return map1<bool>(a, [](float p) { return isnan(p); }); }
template <int N> inline vec<bool, N> isinf(const vec<float, N>& a) { return map1<bool>(a, [](float p) { return isinf(p); }); }
template <int N> inline vec<bool, N> isfinite(const vec<float, N>& a) { return map1<bool>(a, [](float p) { return isfinite(p); }); }
template <int N> inline vec<float, N> min(const vec<float, N>& a, const vec<float, N>& b) { return map2<float>(a, b, [](float p, float q) { return min(p, q); }); }
template <int N> inline vec<float, N> max(const vec<float, N>& a, const vec<float, N>& b) { return map2<float>(a, b, [](float p, float q) { return max(p, q); }); }
template <int N> inline vec<float, N> lerp(const vec<float, N>& a, const vec<float, N>& b, float t) { return a + t * (b - a); }
template <int N> inline vec<float, N> pow(const vec<float, N>& a, float y) { return map1<float>(a, [y](float p) { return pow(p, y); }); }

inline float dot(const float2& a, const float2& b) { return a.x * b.x + a.y * b.y; }
inline float dot(const float3& a, const float3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float dot(const float4& a, const float4& b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }
inline float3 cross(const float3& a, const float3& b) { return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
template <int N> inline float length(const vec<float, N>& a) { return ::sqrtf(dot(a, a)); }
inline float length(float x) { return ::fabsf(x); }            // HLSL length() of a scalar (include/Hit.hlsl:310)
template <int N> inline float distance(const vec<float, N>& a, const vec<float, N>& b) { return length(a - b); }
template <int N> inline vec<float, N> normalize(const vec<float, N>& a) { return a * rsqrt(dot(a, a)); }
inline float3 reflect(const float3& i, const float3& n) { return i - (2.0f * dot(n, i)) * n; }
// half vectors: HLSL evaluates length()/dot() of half operands in half precision (every product and sum rounded to binary16)
inline half dot(const half3& a, const half3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline half length(const half3& a) { return half(::sqrtf(dot(a, a).v)); }
// mixed float3 / half3 arguments promote to float3
inline float dot(const float3& a, const half3& b) { return dot(a, float3(b)); }
inline float dot(const half3& a, const float3& b) { return dot(float3(a), b); }

// ------------------------------------------------------------------------------------------------ matrices
// float4x4 = the 64 bytes the host wrote (an XMMATRIX / glm matrix), read with HLSL's default column-major packing
struct float4x4 {
    float m[16];
    float4 operator[](int r) const { return float4(m[r], m[4 + r], m[8 + r], m[12 + r]); }
};
typedef float4x4 matrix;
inline float4 mul(const float4x4& M, const float4& v) {
    const orc::f4 r = orc::mul44(M.m, v.x, v.y, v.z, v.w);
    return float4(r.x, r.y, r.z, r.w);
}

// row-vector form (a `row_major` parameter: element (r, c) of the same 64 bytes is m[4 r + c])
inline float4 mul(const float4& v, const float4x4& M) {
    float4 r;
    for (int c = 0; c < 4; c++) r[c] = ((v.x * M.m[c] + v.y * M.m[4 + c]) + v.z * M.m[8 + c]) + v.w * M.m[12 + c];
    return r;
}

// ------------------------------------------------------------------------------------------------ resources
template <class T> struct StructuredBuffer {
    const T* data = nullptr; uint64_t count = 0;
    template <class I> T operator[](I i) const {
        const uint32_t k = (uint32_t)i;
        if (k < count) return data[k];
        T z; memset((void*)&z, 0, sizeof z); return z;          // robust buffer access: out-of-bounds reads return zeros
    }
};
template <class T> struct RWStructuredBuffer {
    T* data = nullptr; uint64_t count = 0; mutable T sink;
    template <class I> T& operator[](I i) const {
        const uint32_t k = (uint32_t)i;
        if (k < count) return data[k];
        memset((void*)&sink, 0, sizeof sink); return sink;       // out-of-bounds writes are dropped, reads return zeros
    }
};
template <class T> struct RWTexture2D {
    T* data = nullptr; uint w = 0, h = 0; mutable T sink;
    T& operator[](const uint2& p) const { if (p.x < w && p.y < h) return data[(size_t)p.y * w + p.x]; sink = T(); return sink; }
};
template <class T> struct RWTexture2DArray {
    T* data = nullptr; uint w = 0, h = 0, layers = 0; mutable T sink;
    T& operator[](const uint3& p) const {
        if (p.x < w && p.y < h && p.z < layers) return data[((size_t)p.z * h + p.y) * w + p.x];
        sink = T(); return sink;
    }
};
struct RaytracingAccelerationStructure {};
struct RayDesc { float3 Origin; float TMin; float3 Direction; float TMax; };
enum { RAY_FLAG_NONE = 0 };

// ------------------------------------------------------------------------------------------------ DXR system values (set by the harness)
struct DxrState {
    uint3 launch_index, launch_dims;
    uint instance_id = 0, primitive_index = 0; float ray_t = 0; float3 world_origin, world_direction;
    // TraceRay back-ends: fill the payload memory by running the reference's hit / miss shaders (ref_harness.cpp)
    void (*trace_closest)(const RayDesc& ray, void* payload) = nullptr;
    void (*trace_shadow)(const RayDesc& ray, void* payload) = nullptr;
};
extern thread_local DxrState g_dxr;
inline uint3 DispatchRaysIndex() { return g_dxr.launch_index; }
inline uint3 DispatchRaysDimensions() { return g_dxr.launch_dims; }
inline uint InstanceID() { return g_dxr.instance_id; }
inline uint PrimitiveIndex() { return g_dxr.primitive_index; }
inline float RayTCurrent() { return g_dxr.ray_t; }
inline float3 WorldRayOrigin() { return g_dxr.world_origin; }
inline float3 WorldRayDirection() { return g_dxr.world_direction; }
// hit group offset 0 / miss 0 = {ClosestHit, Miss} on a HitInfo payload; offset 1 / miss 1 = {ShadowClosestHit, ShadowMiss}
// (SBT layout of rdn/Renderer.cpp:1592-1658)
template <class Payload>
inline void TraceRay(const RaytracingAccelerationStructure&, uint flags, uint mask, uint hit_group_offset, uint hit_group_stride, uint miss_index,
                     const RayDesc& ray, Payload& payload) {
    (void)flags; (void)mask; (void)hit_group_stride; (void)miss_index;
    if (hit_group_offset == 0) g_dxr.trace_closest(ray, &payload);
    else g_dxr.trace_shadow(ray, &payload);
}

}  // namespace hlsl
