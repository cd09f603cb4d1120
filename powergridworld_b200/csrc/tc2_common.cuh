// Device helpers shared by the tcgen05 split-FP16 power-flow kernels (powerflow_tc2.cu: 128-env
// tiles, step_fused.cu: 32-env tiles fused with the component steps): UMMA descriptors, MMA
// issue / commit, TMEM loads, FP16 splitting, the load characteristic in float32 and float64.
#pragma once
#include <cuda_fp16.h>

#include "internal.cuh"
#include "tma.cuh"

namespace pgw {

constexpr int T2_M = 128;
// resident CTAs per SM
constexpr int t2_ctas_per_sm(int nch) { return nch <= 4 ? 2 : 1; }
// A CTA is G groups of 128 threads (one thread per env row and group); group g owns the chunks
// g, g + G, ...  Wide feeders use G = 4 and one CTA per SM (shared memory is full anyway); small
// ones (NCH <= 4) use G = 2 with two CTAs per SM, so that every thread has a chunk to work on.

__device__ __forceinline__ uint64_t t2_smem_desc(uint32_t saddr, uint32_t sbo) {
  // cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
  // layout_type = SWIZZLE_NONE [61,64)
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(128u >> 4) << 16) |
         ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ uint32_t t2_idesc_f16(int n) {
  // cute::UMMA::InstrDescriptor: c_format=F32 [4,6), a/b_format=F16 (0) [7,10)/[10,13),
  // a/b K-major, n_dim = N>>3 [17,23), m_dim = M>>4 [24,29)
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(T2_M >> 4) << 24);
}

__device__ __forceinline__ void t2_umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void t2_commit(uint64_t* mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(mbar))
               : "memory");
}

__device__ __forceinline__ void t2_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void t2_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
      "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
      "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
      "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
      "r"(__float_as_uint(v[15]))
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}

// (a, b) -> packed FP16 pair {lo half = a, hi half = b}, round to nearest, saturating
__device__ __forceinline__ uint32_t t2_pack(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(mbar)) : "memory");
}

__device__ __forceinline__ float t2_rsqrt(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ bool t2_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.b32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// hi / lo FP16 images of 8 scaled values as two 16-byte vectors
__device__ __forceinline__ void t2_split8(const float (&x)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    h[q] = t2_pack(x[2 * q], x[2 * q + 1]);
    const float2 back = __half22float2(*reinterpret_cast<const __half2*>(&h[q]));
    l[q] = t2_pack(x[2 * q] - back.x, x[2 * q + 1] - back.y);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// Controllable kW of a load that several agents share (multiagent_env.py:171-181: summed per
// load name in agent order).  Rare, so kept out of line: the prologue is unrolled 24 times.
static __device__ __noinline__ double t2_shared_load_kw(const double* agent_p, const int32_t* lidx, int q0,
                                                 int q1, int E, int e) {
  double s = 0.0;
#pragma unroll 1
  for (int q = q0; q < q1; q += 4) {                   // four loads in flight, summed in order
    double v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = q + u < q1 ? agent_p[(size_t)lidx[q + u] * E + e] : 0.0;
#pragma unroll
    for (int u = 0; u < 4; ++u) s += v[u];
  }
  return s;
}

// Per-branch constants (shared memory, read as warp-uniform broadcasts):
//   cst[k] = {Re u0, Im u0, vlo^2, vhi^2}: current = conj(s) u / clamp(|u|^2, vlo^2, vhi^2), i.e.
//            constant PQ inside the band and constant Z outside (OpenDSS model 1); a constant-Z
//            load (model 2) is the degenerate band [1, 1]; a constant-current load (model 5) has
//            the open band [1e-30, 3e38] and
//   gh[k]  = {g, h}: k = r (r g + h) with r = rsqrt(clamp): (1, 0) -> 1/clamp, (0, 1) -> 1/|u|.
template <bool ANY_M5>
__device__ __forceinline__ void t2_current(float4 c, float2 gh, float dr, float di, float ds,
                                           float sr, float si, float& x, float& y) {
  const float ur = fmaf(dr, ds, c.x), ui = fmaf(di, ds, c.y);
  const float m2 = fmaf(ur, ur, ui * ui);
  const float r = t2_rsqrt(fminf(fmaxf(m2, c.z), c.w));
  const float kf = ANY_M5 ? r * fmaf(r, gh.x, gh.y) : r * r;
  const float tr = ur * kf, ti = ui * kf;
  x = fmaf(sr, tr, si * ti);                           // conj(s) u k
  y = fmaf(sr, ti, -(si * tr));
}

// STANDALONE: the solve of pgw_pf_solve (total kW / kvar per load given per env) instead of the
// step / reset solve (base load of the event + the agents' powers).
// OCC: resident CTAs per SM the kernel is compiled for.  Small feeders (NCH <= 2, one chunk per
// thread) also come in a 64-register build that runs four CTAs per SM: slower for a single wave
// of tiles (spills, latency) but ~12 % faster once the batch is several waves deep.
// The same for the branch pair (2q, 2q + 1) in packed f32x2 arithmetic (FFMA2 / FMUL2): half the
// issue slots for the multiply-add part.  Lane-wise identical to t2_current.
template <bool ANY_M5>
__device__ __forceinline__ void t2_current2(float4 a, float4 b, float4 c, float2 dr, float2 di, float ds,
                                            float2 sr, float2 si, float2& x, float2& y) {
  const float2 ds2 = make_float2(ds, ds);
  const float2 ur = __ffma2_rn(dr, ds2, make_float2(a.x, a.y));
  const float2 ui = __ffma2_rn(di, ds2, make_float2(a.z, a.w));
  const float2 m2 = __ffma2_rn(ur, ur, __fmul2_rn(ui, ui));
  const float2 r = make_float2(t2_rsqrt(fminf(fmaxf(m2.x, b.x), b.z)),
                               t2_rsqrt(fminf(fmaxf(m2.y, b.y), b.w)));
  const float2 kf = ANY_M5 ? __fmul2_rn(r, __ffma2_rn(r, make_float2(c.x, c.y), make_float2(c.z, c.w)))
                           : __fmul2_rn(r, r);
  const float2 tr = __fmul2_rn(ur, kf), ti = __fmul2_rn(ui, kf);
  x = __ffma2_rn(sr, tr, __fmul2_rn(si, ti));          // conj(s) u k
  const float2 m = __fmul2_rn(si, tr);
  y = __ffma2_rn(sr, ti, make_float2(-m.x, -m.y));
}

// scalar views of the pair-packed constants (cold paths)
__device__ __forceinline__ float4 t2_cst(const Tc2Consts& kc, int k) {
  const float* a = &kc.pa[k >> 1].x;
  const float* b = &kc.pb[k >> 1].x;
  const int o = k & 1;
  return make_float4(a[o], a[2 + o], b[o], b[2 + o]);
}
__device__ __forceinline__ float2 t2_gh(const Tc2Consts& kc, int k) {
  const float* c = &kc.pc[k >> 1].x;
  return make_float2(c[k & 1], c[2 + (k & 1)]);
}

// The load characteristic in float32 with correctly rounded reciprocals (the hot loop's
// rsqrt.approx is good to 2^-22.9: squared, 2.6e-7 relative on a current) -- currents SCALED by
// xscale like t2_current's.
template <bool ANY_M5>
__device__ __forceinline__ void t2_current32(float4 c, float2 gh, float dr, float di, float ds, float sr,
                                             float si, float& x, float& y) {
  const float ur = fmaf(dr, ds, c.x), ui = fmaf(di, ds, c.y);
  const float m2 = fmaf(ur, ur, ui * ui);
  const float cl = fminf(fmaxf(m2, c.z), c.w);
  float kf = __frcp_rn(cl);
  if (ANY_M5) kf = fmaf(gh.y, __frsqrt_rn(cl), gh.x * kf);
  const float tr = ur * kf, ti = ui * kf;
  x = fmaf(sr, tr, si * ti);                           // conj(s) u k
  y = fmaf(sr, ti, -(si * tr));
}
__device__ __forceinline__ void t2_cmac_sub32(float2& acc, float2 z, float2 i) {
  acc.x = fmaf(-z.x, i.x, acc.x);
  acc.x = fmaf(z.y, i.y, acc.x);
  acc.y = fmaf(-z.x, i.y, acc.y);
  acc.y = fmaf(-z.y, i.x, acc.y);
}

// ---- FP64 polish (POLISH instantiations): the load characteristic and one row of the sweep
// u <- u0 - Zbb i(u) in float64, same model semantics as branch_current of powerflow.cu
// (OpenDSS model 1 = constant PQ inside [vmin, vmax], constant Z outside; 2 = constant Z;
// 5 = constant current magnitude).
__device__ __forceinline__ double2 t2_current64(int model, double2 s, double2 u, double vlo2,
                                                double vhi2) {
  const double m2 = u.x * u.x + u.y * u.y;
  double k;
  if (model == 5) {
    k = m2 > 0.0 ? rsqrt(m2) : 0.0;
  } else {                                               // model 2: the band is [1, 1]
    // 1 / c from the float32 reciprocal and two Newton steps (relative error ~1e-21 before
    // rounding): the voltages are O(1), nothing here can overflow or go subnormal
    const double c = fmin(fmax(m2, vlo2), vhi2);
    double x = (double)__frcp_rn((float)c);
    x = fma(x, fma(-c, x, 1.0), x);
    k = fma(x, fma(-c, x, 1.0), x);
  }
  const double tr = s.x * u.x + s.y * u.y, ti = s.x * u.y - s.y * u.x;   // conj(s) u
  return make_double2(tr * k, ti * k);
}
__device__ __forceinline__ void t2_cmac_sub64(double2& acc, double2 z, double2 i) {
  acc.x = fma(-z.x, i.x, acc.x);
  acc.x = fma(z.y, i.y, acc.x);
  acc.y = fma(-z.x, i.y, acc.y);
  acc.y = fma(-z.y, i.x, acc.y);
}

// One float64 sweep over the 8 branches of chunk c of env row `row`:
// u_k <- u0_k - sum_j Zbb[k][j] i_j.  Z and u0 are warp-uniform constant-bank reads, the env's
// currents come from shared memory, [branch][env row].  The loop over j stays rolled: the code
// runs once per tile, and straight-line code would be paid for in instruction fetches.
__device__ __forceinline__ void t2_sweep64(const Tc2Polish& kp, int nb, int c, const double2* sI,
                                           int row, double2 (&u)[8]) {
  double2 acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = kp.u0[8 * c + j];
#pragma unroll 2
  for (int jj = 0; jj < nb; ++jj) {
    const double2 ij = sI[(size_t)jj * T2_M + row];
    const double2* z = kp.zT + jj * 16 + 8 * c;
#pragma unroll
    for (int j = 0; j < 8; ++j) t2_cmac_sub64(acc[j], z[j], ij);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (8 * c + j < nb) u[j] = acc[j];
}

}  // namespace pgw
