// Small fixed-size vector / quaternion / spatial-algebra types used by the per-env pipeline.
// Everything is by-value and fully inlined so that it lives in registers.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define GG_HD __host__ __device__ __forceinline__
#else
#define GG_HD inline
#endif

namespace gg {

GG_HD int popc32(unsigned v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}

struct V3 { float x, y, z; };
GG_HD V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
GG_HD V3 ld3(const float* p) { return mk(p[0], p[1], p[2]); }
GG_HD V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
GG_HD V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
GG_HD V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
GG_HD V3 operator*(float s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
GG_HD V3 operator*(V3 a, float s) { return mk(s * a.x, s * a.y, s * a.z); }
GG_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
GG_HD V3 cross(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
GG_HD float comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

struct Q4 { float w, x, y, z; };
GG_HD Q4 mkq(float w, float x, float y, float z) { Q4 r; r.w = w; r.x = x; r.y = y; r.z = z; return r; }
GG_HD Q4 ldq(const float* p) { return mkq(p[0], p[1], p[2], p[3]); }
GG_HD Q4 qmul(Q4 a, Q4 b) {
    return mkq(a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
               a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w);
}
// q (x) rotation by `ang` about local z, given the half-angle sine / cosine
GG_HD Q4 qmul_z(Q4 a, float c, float s) { return mkq(a.w * c - a.z * s, a.x * c + a.y * s, a.y * c - a.x * s, a.z * c + a.w * s); }

struct M3 { V3 c0, c1, c2; };   // columns
GG_HD M3 qmat(Q4 q) {
    M3 R;
    R.c0 = mk(1.f - 2.f * (q.y * q.y + q.z * q.z), 2.f * (q.x * q.y + q.w * q.z), 2.f * (q.x * q.z - q.w * q.y));
    R.c1 = mk(2.f * (q.x * q.y - q.w * q.z), 1.f - 2.f * (q.x * q.x + q.z * q.z), 2.f * (q.y * q.z + q.w * q.x));
    R.c2 = mk(2.f * (q.x * q.z + q.w * q.y), 2.f * (q.y * q.z - q.w * q.x), 1.f - 2.f * (q.x * q.x + q.y * q.y));
    return R;
}
GG_HD V3 mul(const M3& R, V3 v) { return R.c0 * v.x + R.c1 * v.y + R.c2 * v.z; }
GG_HD V3 mulT(const M3& R, V3 v) { return mk(dot(R.c0, v), dot(R.c1, v), dot(R.c2, v)); }
GG_HD M3 mmul(const M3& A, const M3& B) { M3 r; r.c0 = mul(A, B.c0); r.c1 = mul(A, B.c1); r.c2 = mul(A, B.c2); return r; }
GG_HD V3 col(const M3& R, int i) { return i == 0 ? R.c0 : (i == 1 ? R.c1 : R.c2); }

struct S3 { float xx, yy, zz, xy, xz, yz; };
GG_HD V3 mul(const S3& I, V3 v) {
    return mk(I.xx * v.x + I.xy * v.y + I.xz * v.z, I.xy * v.x + I.yy * v.y + I.yz * v.z, I.xz * v.x + I.yz * v.y + I.zz * v.z);
}

// Spatial vectors about the WORLD ORIGIN, world orientation.  Motion: (omega, velocity of the
// body-fixed point at O).  Force: (moment about O, force).
struct SV { V3 a, l; };
GG_HD SV operator+(SV p, SV q) { SV r; r.a = p.a + q.a; r.l = p.l + q.l; return r; }
GG_HD SV operator*(SV p, float s) { SV r; r.a = p.a * s; r.l = p.l * s; return r; }
GG_HD float sdot(const SV& s, const SV& f) { return dot(s.a, f.a) + dot(s.l, f.l); }
GG_HD SV crm(const SV& v, const SV& s) { SV r; r.a = cross(v.a, s.a); r.l = cross(v.a, s.l) + cross(v.l, s.a); return r; }
GG_HD SV crf(const SV& v, const SV& f) { SV r; r.a = cross(v.a, f.a) + cross(v.l, f.l); r.l = cross(v.a, f.l); return r; }

// Spatial inertia about O: mass, first moment h = m c, second moment about O.
struct SI { float m; V3 h; S3 I; };
GG_HD SI operator+(const SI& p, const SI& q) {
    SI r; r.m = p.m + q.m; r.h = p.h + q.h;
    r.I.xx = p.I.xx + q.I.xx; r.I.yy = p.I.yy + q.I.yy; r.I.zz = p.I.zz + q.I.zz;
    r.I.xy = p.I.xy + q.I.xy; r.I.xz = p.I.xz + q.I.xz; r.I.yz = p.I.yz + q.I.yz;
    return r;
}
GG_HD SV apply(const SI& I, const SV& v) { SV r; r.a = mul(I.I, v.a) + cross(I.h, v.l); r.l = I.m * v.l + cross(v.a, I.h); return r; }

// world spatial inertia of a body with pose (R, p), local COM `com`, body-frame inertia i6 = xx yy zz xy xz yz at the COM
GG_HD SI body_inertia(float m, V3 com, const float* i6, const M3& R, V3 p) {
    V3 c = p + mul(R, com);
    V3 t0 = mul(R, mk(i6[0], i6[3], i6[4])), t1 = mul(R, mk(i6[3], i6[1], i6[5])), t2 = mul(R, mk(i6[4], i6[5], i6[2]));
    SI r; r.m = m; r.h = m * c;
    float cc = dot(c, c);
    r.I.xx = t0.x * R.c0.x + t1.x * R.c1.x + t2.x * R.c2.x + m * (cc - c.x * c.x);
    r.I.yy = t0.y * R.c0.y + t1.y * R.c1.y + t2.y * R.c2.y + m * (cc - c.y * c.y);
    r.I.zz = t0.z * R.c0.z + t1.z * R.c1.z + t2.z * R.c2.z + m * (cc - c.z * c.z);
    r.I.xy = t0.x * R.c0.y + t1.x * R.c1.y + t2.x * R.c2.y - m * c.x * c.y;
    r.I.xz = t0.x * R.c0.z + t1.x * R.c1.z + t2.x * R.c2.z - m * c.x * c.z;
    r.I.yz = t0.y * R.c0.z + t1.y * R.c1.z + t2.y * R.c2.z - m * c.y * c.z;
    return r;
}

GG_HD void gg_sincos(float a, float* s, float* c) {
#if defined(__CUDA_ARCH__)
    sincosf(a, s, c);
#else
    *s = sinf(a); *c = cosf(a);
#endif
}

// ---- packed lower-triangular helpers (row-major: idx(i,j) = i(i+1)/2 + j, j <= i) -----------------
#define GG_TRI(i, j) ((i) * ((i) + 1) / 2 + (j))

// fast reciprocal / reciprocal square root (device: MUFU-based, <= 2 ulp; host harness: exact)
GG_HD float gg_rcp(float x) {
#if defined(__CUDA_ARCH__)
    return __frcp_rn(x);
#else
    return 1.0f / x;
#endif
}
GG_HD float gg_div(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdividef(a, b);
#else
    return a / b;
#endif
}
GG_HD float gg_rsqrt(float x) {
#if defined(__CUDA_ARCH__)
    float r = rsqrtf(x);
    return r * (1.5f - 0.5f * x * r * r);      // one Newton step: ~1 ulp
#else
    return 1.0f / sqrtf(x);
#endif
}

// in-place Cholesky A = L L^T on a packed lower matrix, fully unrolled (static indices -> registers).
// dinv receives 1 / L(j,j) so that the triangular solves multiply instead of divide.
template <int N> GG_HD bool chol_packed(float* A, float* dinv) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < N; j++) {
        float s = A[GG_TRI(j, j)];
#pragma unroll
        for (int k = 0; k < j; k++) s -= A[GG_TRI(j, k)] * A[GG_TRI(j, k)];
        ok = ok && (s > 0.f);
        float inv = gg_rsqrt(s);
        dinv[j] = inv;
        A[GG_TRI(j, j)] = s * inv;
#pragma unroll
        for (int i = j + 1; i < N; i++) {
            float t = A[GG_TRI(i, j)];
#pragma unroll
            for (int k = 0; k < j; k++) t -= A[GG_TRI(i, k)] * A[GG_TRI(j, k)];
            A[GG_TRI(i, j)] = t * inv;
        }
    }
    return ok;
}
// solve L L^T x = b (x may alias b)
template <int N> GG_HD void chol_solve_packed(const float* L, const float* dinv, const float* b, float* x) {
    float y[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        float s = b[i];
#pragma unroll
        for (int k = 0; k < i; k++) s -= L[GG_TRI(i, k)] * y[k];
        y[i] = s * dinv[i];
    }
#pragma unroll
    for (int i = N - 1; i >= 0; i--) {
        float s = y[i];
#pragma unroll
        for (int k = i + 1; k < N; k++) s -= L[GG_TRI(k, i)] * x[k];
        x[i] = s * dinv[i];
    }
}
// y = A x for a packed symmetric matrix
template <int N> GG_HD void symv_packed(const float* A, const float* x, float* y) {
#pragma unroll
    for (int i = 0; i < N; i++) {
        float r = 0.f;
#pragma unroll
        for (int j = 0; j < N; j++) r += A[i >= j ? GG_TRI(i, j) : GG_TRI(j, i)] * x[j];
        y[i] = r;
    }
}

// ---- Philox4x32-10 (Salmon et al., Random123) -----------------------------------------------------
GG_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
GG_HD float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; }

}  // namespace gg
