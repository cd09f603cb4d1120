// K6, tensor-core form: the Z-bus fixed point  u <- u0 - Zbb i(u)  as a dense contraction on
// the 5th-generation tensor cores (tcgen05), one CTA per tile of 128 envs.
//
//   D[env, :] = X[env, :] * B^T        M = 128 envs (TMEM lanes), N = 2*16 (Re|Im of du),
//                                       K = 96 = 3 x 32 (split-TF32, see below)
//
// * The env batch sits on the MMA M dimension, so after tcgen05.ld every thread owns the
//   complete voltage-drop vector of ITS env: the nonlinear load characteristic, the
//   convergence test and the convergence mask are thread-local (no shuffles, no reductions).
// * Real-ified complex product: x = [Re i | Im i] (32 reals), B = -[[Zr, -Zi], [Zi, Zr]].
// * kind::tf32 keeps 10 mantissa bits, so operands are split  x = x_hi + x_lo,  B = B_hi + B_lo
//   and the three significant products are one accumulation chain over a tripled K:
//   A' = [x_hi | x_lo | x_hi],  B' = [B_hi ; B_hi ; B_lo]  (12 MMAs of K = 8), FP32 accumulate in
//   TMEM.  Only the voltage DROP (<= ~0.1 p.u.) goes through the tensor core; u = u0 + drop.
// * A' is produced by the epilogue threads straight into shared memory in the canonical
//   K-major no-swizzle UMMA layout (8-row x 16-byte core matrices); B' images are prepared
//   on the host and staged once per CTA with a TMA bulk copy.
// * After convergence one more MMA chain with B2' (N = 2*nn) expands to all node voltages.
//
// Same inputs/outputs as pf_fixed_point_kernel (powerflow.cu); voltages agree with it to
// ~1e-7 p.u. (FP32 epilogue), inside the 1e-4 p.u. tolerance of the reference's solver.
#include "internal.cuh"
#include "tma.cuh"

namespace pgw {

constexpr int TC_M = 128;
constexpr uint32_t TC_LBO = 128;                       // between the two 16-byte K chunks of one MMA
constexpr uint32_t TC_SBO = (kTcK3 / 4) * 128;         // between 8-row groups (3072 B)
constexpr uint32_t TC_A_BYTES = TC_M * kTcK3 * 4;      // 49152
constexpr uint32_t TC_COLS = 128;                      // TMEM columns: D1 at 0..31, D2 at 32..

__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr) {
  // cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
  // layout_type = SWIZZLE_NONE [61,64)
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(TC_LBO >> 4) << 16) |
         ((uint64_t)(TC_SBO >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ uint32_t umma_idesc_tf32(int n) {
  // cute::UMMA::InstrDescriptor: c_format=F32 [4,6), a/b_format=TF32 [7,10)/[10,13),
  // a/b K-major, n_dim = N>>3 [17,23), m_dim = M>>4 [24,29)
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(TC_M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(mbar))
               : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
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

__device__ __forceinline__ float tf32_hi(float x) {
  return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
}

__global__ void __launch_bounds__(TC_M) pf_tc_kernel(const PfParams p) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  __shared__ __align__(8) uint64_t mbar_tma, mbar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int hdr = 2 + 2 * p.nl;

  unsigned char* sA = tc_smem;                                   // A' tile, 48 kB
  unsigned char* sT = tc_smem + TC_A_BYTES;                      // staged blob: B1' | B2' | tables
  double* drow = reinterpret_cast<double*>(sT + p.tc_blob_bytes);

  const int event = p.event_mode == 0 ? 0 : (*p.clock + 1);
  if (tid == 0) {
    mbar_init(&mbar_tma, 1);
    mbar_init(&mbar_mma, 1);
    mbar_expect_tx(&mbar_tma, (uint32_t)p.tc_blob_bytes + (uint32_t)hdr * 8u);
    tma_bulk_g2s(sT, p.tc_blob, (uint32_t)p.tc_blob_bytes, &mbar_tma);
    tma_bulk_g2s(drow, p.dtab + (size_t)event * p.dstride, (uint32_t)hdr * 8u, &mbar_tma);
  }
  if (warp == 0) {                                               // one warp owns TMEM alloc / free
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_s)),
                 "r"(TC_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  mbar_wait(&mbar_tma, 0);

  const float2* u0f = reinterpret_cast<const float2*>(sT + p.tc_off_u0);
  const float2* wf = reinterpret_cast<const float2*>(sT + p.tc_off_w);
  const double* share = reinterpret_cast<const double*>(sT + p.tc_off_share);
  const float* vlo = reinterpret_cast<const float*>(sT + p.tc_off_vlo);
  const float* vhi = reinterpret_cast<const float*>(sT + p.tc_off_vhi);
  const int32_t* bload = reinterpret_cast<const int32_t*>(sT + p.tc_off_bload);
  const int32_t* bmodel = reinterpret_cast<const int32_t*>(sT + p.tc_off_bmodel);
  const int32_t* aslot = reinterpret_cast<const int32_t*>(sT + p.tc_off_slot);
  const int32_t* anode = reinterpret_cast<const int32_t*>(sT + p.tc_off_node);
  const double* base_kw = drow + 2;
  const double* base_kvar = drow + 2 + p.nl;

  const uint32_t a_addr = smem_u32(sA);
  const uint32_t b1_addr = smem_u32(sT);
  const uint32_t b2_addr = smem_u32(sT + p.tc_off_b2);
  const uint32_t idesc1 = umma_idesc_tf32(32), idesc2 = umma_idesc_tf32(p.tc_n2);
  unsigned char* a_row = sA + (size_t)(tid >> 3) * TC_SBO + (size_t)(tid & 7) * 16;
  const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);   // this warp's 32 TMEM lanes
  uint32_t mma_phase = 0;

  const int tiles = (p.E + TC_M - 1) / TC_M;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int e_raw = tile * TC_M + tid;
    const bool valid = e_raw < p.E;
    const int e = valid ? e_raw : p.E - 1;

    // ---- per-branch nominal power (p.u. on 1 MVA) and warm start, in registers
    float sr[kTcNb], si[kTcNb], dr[kTcNb], di[kTcNb];
#pragma unroll
    for (int k = 0; k < kTcNb; ++k) {
      sr[k] = si[k] = dr[k] = di[k] = 0.f;
      if (k < p.nb) {
        const int l = bload[k];
        double kw, kvar;
        if (p.load_kw != nullptr) {
          kw = p.load_kw[(size_t)l * p.E + e];
          kvar = p.load_kvar[(size_t)l * p.E + e];
        } else {
          kw = base_kw[l];
          kvar = base_kvar[l];
          if (p.agent_p != nullptr) {                 // multiagent_env.py:171-181, opendss.py:128
            double ctrl = 0.0;
            bool any = false;
            for (int a = 0; a < p.A; ++a)
              if (aslot[a] == l) {
                const double pa = p.agent_p[(size_t)a * p.E + e];
                ctrl = any ? ctrl + pa : pa;
                any = true;
              }
            if (any) kw += ctrl;
          }
        }
        const double sh = share[k] * 1e-3;
        sr[k] = (float)(kw * sh);
        si[k] = (float)(kvar * sh);
        if (p.warm_start) {
          const double2 up = p.u_state[(size_t)k * p.E + e];
          dr[k] = (float)(up.x - (double)u0f[k].x);
          di[k] = (float)(up.y - (double)u0f[k].y);
        }
      }
    }

    int it = 0;
    bool conv = !valid, conv_ok = true;
    while (true) {
      // ---- branch currents at u = u0 + drop, split hi/lo, written as this env's row of A'
      float xr[kTcNb], xi[kTcNb];
#pragma unroll
      for (int k = 0; k < kTcNb; ++k) {
        const float ur = u0f[k].x + dr[k], ui = u0f[k].y + di[k];
        const float m2 = ur * ur + ui * ui;
        const int model = bmodel[k];
        float kf;
        if (model == 2) kf = 1.f;
        else if (model == 5) kf = m2 > 0.f ? rsqrtf(m2) : 0.f;
        else if (m2 <= vlo[k] * vlo[k]) kf = __frcp_rn(vlo[k] * vlo[k]);
        else if (m2 > vhi[k] * vhi[k]) kf = __frcp_rn(vhi[k] * vhi[k]);
        else kf = __frcp_rn(m2);
        xr[k] = (sr[k] * ur + si[k] * ui) * kf;       // conj(s) u k
        xi[k] = (sr[k] * ui - si[k] * ur) * kf;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 full, hi, lo;
        const float* src = q < 4 ? &xr[4 * q] : &xi[4 * (q - 4)];
        full = make_float4(src[0], src[1], src[2], src[3]);
        hi = make_float4(tf32_hi(full.x), tf32_hi(full.y), tf32_hi(full.z), tf32_hi(full.w));
        lo = make_float4(full.x - hi.x, full.y - hi.y, full.z - hi.z, full.w - hi.w);
        *reinterpret_cast<float4*>(a_row + q * 128) = hi;            // K  0..31  x_hi
        *reinterpret_cast<float4*>(a_row + 1024 + q * 128) = lo;     // K 32..63  x_lo
        *reinterpret_cast<float4*>(a_row + 2048 + q * 128) = hi;     // K 64..95  x_hi
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // st.shared -> tensor core
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      const int all_done = __syncthreads_and(conv ? 1 : 0);
      if (all_done) break;                                           // A' = currents at final u

      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int kk = 0; kk < kTcK3 / 8; ++kk)
          umma_tf32(tmem, umma_smem_desc(a_addr + kk * 256), umma_smem_desc(b1_addr + kk * 256),
                    idesc1, kk > 0 ? 1u : 0u);
        umma_commit(&mbar_mma);
      }
      mbar_wait(&mbar_mma, mma_phase);
      mma_phase ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

      float nr[16], ni[16];
      tmem_ld16(t_lane + 0, nr);                                     // Re(drop_k)
      tmem_ld16(t_lane + 16, ni);                                    // Im(drop_k)
      if (!conv) {
        float d = 0.f;
#pragma unroll
        for (int k = 0; k < kTcNb; ++k) {
          d = fmaxf(d, fmaxf(fabsf(nr[k] - dr[k]), fabsf(ni[k] - di[k])));
          dr[k] = nr[k];
          di[k] = ni[k];
        }
        ++it;
        conv_ok = d < p.tc_tol;
        conv = conv_ok || it >= p.max_iter;
      }
    }

    // ---- expansion to all node voltages: D2 = X * B2'^T into TMEM columns 32..
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int kk = 0; kk < kTcK3 / 8; ++kk)
        umma_tf32(tmem + 32, umma_smem_desc(a_addr + kk * 256), umma_smem_desc(b2_addr + kk * 256),
                  idesc2, kk > 0 ? 1u : 0u);
      umma_commit(&mbar_mma);
    }
    mbar_wait(&mbar_mma, mma_phase);
    mma_phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    float vmn = 3.0e38f, vmx = -3.0e38f;
    for (int c = 0; c < p.tc_n2 / 16; ++c) {                         // 8 nodes per chunk: (Re, Im) pairs
      float v[16];
      tmem_ld16(t_lane + 32 + 16 * c, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = c * 8 + j;
        if (n < p.nn) {
          const float vr = wf[n].x + v[2 * j], vi = wf[n].y + v[2 * j + 1];
          const float mag = __fsqrt_rn(vr * vr + vi * vi);
          vmn = fminf(vmn, mag);
          vmx = fmaxf(vmx, mag);
          if (valid) p.vmag[(size_t)n * p.E + e] = (double)mag;
        }
      }
    }

    if (valid) {
#pragma unroll
      for (int k = 0; k < kTcNb; ++k)
        if (k < p.nb)
          p.u_state[(size_t)k * p.E + e] =
              make_double2((double)u0f[k].x + (double)dr[k], (double)u0f[k].y + (double)di[k]);
      p.vmin[e] = (double)vmn;
      p.vmax[e] = (double)vmx;
      p.iters[e] = conv_ok ? it : -it;
      double pen_share = 0.0, viol = 0.0;
      if (p.punit != 0.0) {
        const double v = p.vmag[(size_t)p.penalty_node * p.E + e];   // this thread's own store
        viol = fmax(0.0, fmax(p.pvlo - v, v - p.pvhi));
        pen_share = (viol * p.punit) / (double)p.A;
      }
      p.viol[e] = viol;
      for (int a = 0; a < p.A; ++a) {
        const int node = anode[a];
        const size_t ae = (size_t)a * p.E + e;
        p.vbus[ae] = node >= 0 ? p.vmag[(size_t)node * p.E + e] : 1.0;
        if (p.event_mode != 0) {
          const double r = p.rew[ae] - pen_share;
          p.rew[ae] = r;
          p.rew_copy[ae] = r;
          p.ep_ret[ae] += r;
        }
      }
    }
    // all TMEM reads of this tile are complete (tcgen05.wait::ld) before the next tile's MMAs
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }

  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_COLS)
                 : "memory");
  if (p.advance_clock) publish_clock_last_cta(p.ticket, p.clock, event, gridDim.x);
}

cudaError_t launch_powerflow_tc(const PfParams& p, cudaStream_t s) {
  const int tiles = (p.E + TC_M - 1) / TC_M;
  int grid = tiles < 148 * 2 ? tiles : 148 * 2;
  if (grid < 1) grid = 1;
  const size_t smem = (size_t)TC_A_BYTES + (size_t)p.tc_blob_bytes + (size_t)(2 + 2 * p.nl) * 8 + 16;
  cudaError_t err = cudaFuncSetAttribute(pf_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
  if (err != cudaSuccess) return err;
  pf_tc_kernel<<<grid, TC_M, smem, s>>>(p);
  return cudaGetLastError();
}

}  // namespace pgw
