// se3_device.cuh -- forward-mode dual numbers and the 6D -> R evaluation shared by se3.cu and relax.cu.
#pragma once
#include "common.cuh"

namespace reart {

// ----------------------------------------------------------------------------- forward-mode duals
template <int ND>
struct Dual {
    float v;
    float d[ND];
    __device__ Dual() {}
    __device__ explicit Dual(float c) : v(c) {
#pragma unroll
        for (int i = 0; i < ND; ++i) d[i] = 0.f;
    }
    __device__ static Dual var(float c, int k) {
        Dual r(c);
        r.d[k] = 1.f;
        return r;
    }
};
template <int ND>
__device__ __forceinline__ Dual<ND> operator+(const Dual<ND>& a, const Dual<ND>& b) {
    Dual<ND> r; r.v = a.v + b.v;
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] + b.d[i];
    return r;
}
template <int ND>
__device__ __forceinline__ Dual<ND> operator-(const Dual<ND>& a, const Dual<ND>& b) {
    Dual<ND> r; r.v = a.v - b.v;
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] - b.d[i];
    return r;
}
template <int ND>
__device__ __forceinline__ Dual<ND> operator-(const Dual<ND>& a) {
    Dual<ND> r; r.v = -a.v;
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = -a.d[i];
    return r;
}
template <int ND>
__device__ __forceinline__ Dual<ND> operator*(const Dual<ND>& a, const Dual<ND>& b) {
    Dual<ND> r; r.v = a.v * b.v;
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
    return r;
}
template <int ND>
__device__ __forceinline__ Dual<ND> operator*(const Dual<ND>& a, float s) {
    Dual<ND> r; r.v = a.v * s;
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] * s;
    return r;
}
template <int ND>
__device__ __forceinline__ Dual<ND> operator/(const Dual<ND>& a, const Dual<ND>& b) {
    Dual<ND> r; r.v = a.v / b.v;
    const float inv = 1.0f / b.v;
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
    return r;
}
template <int ND>
__device__ __forceinline__ Dual<ND> dsin(const Dual<ND>& a) {
    Dual<ND> r; r.v = sinf(a.v);
    const float c = cosf(a.v);
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = c * a.d[i];
    return r;
}
template <int ND>
__device__ __forceinline__ Dual<ND> dcos(const Dual<ND>& a) {
    Dual<ND> r; r.v = cosf(a.v);
    const float s = -sinf(a.v);
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = s * a.d[i];
    return r;
}
template <int ND>
__device__ __forceinline__ Dual<ND> dsqrt(const Dual<ND>& a) {
    Dual<ND> r; r.v = sqrtf(a.v);
    const float k = r.v > 0.f ? 0.5f / r.v : 0.f;            // torch: d sqrt / norm at 0 -> 0 subgradient
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = k * a.d[i];
    return r;
}
// torch.clamp(x, min=lo): passes the gradient where x >= lo, zero below
template <int ND>
__device__ __forceinline__ Dual<ND> dclamp_min(const Dual<ND>& a, float lo) {
    if (a.v >= lo) return a;
    return Dual<ND>(lo);
}
// plain-float overloads so the same templated code serves the forward-only kernels
__device__ __forceinline__ float dsin(float a) { return sinf(a); }
__device__ __forceinline__ float dcos(float a) { return cosf(a); }
__device__ __forceinline__ float dsqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ float dclamp_min(float a, float lo) { return a >= lo ? a : lo; }

template <typename S> __device__ __forceinline__ S lit(float c);
template <> __device__ __forceinline__ float lit<float>(float c) { return c; }
template <> __device__ __forceinline__ Dual<6> lit<Dual<6>>(float c) { return Dual<6>(c); }
template <> __device__ __forceinline__ Dual<8> lit<Dual<8>>(float c) { return Dual<8>(c); }
__device__ __forceinline__ float val(float a) { return a; }
template <int ND> __device__ __forceinline__ float val(const Dual<ND>& a) { return a.v; }

// ----------------------------------------------------------------------------- 6D -> R
// F.normalize(v, eps=1e-12): v / max(|v|, eps)
template <typename S>
__device__ __forceinline__ void normalize3(const S* v, S* o) {
    S n = dsqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    S den = dclamp_min(n, 1e-12f);
    o[0] = v[0] / den; o[1] = v[1] / den; o[2] = v[2] / den;
}
template <typename S>
__device__ __forceinline__ void rot6d_eval(const S* d6, S* R) {
    S b1[3], b2[3], u[3];
    normalize3(d6, b1);
    S dot = b1[0] * d6[3] + b1[1] * d6[4] + b1[2] * d6[5];
    for (int k = 0; k < 3; ++k) u[k] = d6[3 + k] - dot * b1[k];
    normalize3(u, b2);
    R[0] = b1[0]; R[1] = b1[1]; R[2] = b1[2];
    R[3] = b2[0]; R[4] = b2[1]; R[5] = b2[2];
    R[6] = b1[1] * b2[2] - b1[2] * b2[1];
    R[7] = b1[2] * b2[0] - b1[0] * b2[2];
    R[8] = b1[0] * b2[1] - b1[1] * b2[0];
}


}  // namespace reart
