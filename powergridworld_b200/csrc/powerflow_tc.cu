// K6, tensor-core form: the Z-bus fixed point  u <- u0 - Zbb i(u)  as a dense contraction on
// the 5th-generation tensor cores (tcgen05), one CTA per tile of 128 envs.
//
//   D[env, :] = X[env, :] * B^T        M = 128 envs (TMEM lanes), N = 2*16 (Re|Im of du),
//                                       K = 96 = 3 x 32 (split-TF32, see below)
//
// * The env batch sits on the MMA M dimension: TMEM lane r = env r of the tile.  The four
//   warps sharing a lane quadrant split the D columns (4 branches each), so the nonlinear
//   load characteristic is thread-local and the per-env convergence test is one shared-memory
//   max over four partials; converged envs freeze their state (convergence mask).
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

// x4 variant for the 4-branch column slices of the iteration epilogue
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}

// 512 threads = 4 warp groups x 128 envs.  Thread (g, r) owns branches 4g..4g+3 of env row r:
// the four warps that share a TMEM lane quadrant (w, w+4, w+8, w+12) split the columns, which
// quarters the serial epilogue chain and gives the SM 16-32 warps to hide latency with.
constexpr int TC_THREADS = 512;
constexpr int TC_BPT = kTcNb / 4;                      // branches per thread

__global__ void __launch_bounds__(TC_THREADS, 2) pf_tc_kernel(const PfParams p) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  __shared__ __align__(8) uint64_t mbar_tma, mbar_mma;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_dpart[4][TC_M];                   // per-group partial max |d drop|
  __shared__ float s_vmn[4][TC_M], s_vmx[4][TC_M];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & (TC_M - 1);                    // env row in the tile = TMEM lane
  const int grp = tid >> 7;                            // branch group 0..3
  const int hdr = 2 + 2 * p.nl;

  unsigned char* sA = tc_smem;                                   // A' tile, 48 kB
  unsigned char* sT = tc_smem + TC_A_BYTES;                      // staged blob: B1' | B2' | tables
  double* drow = reinterpret_cast<double*>(sT + p.tc_blob_bytes);

  const int clk = p.event_mode == 0 ? -1 : *p.clock;
  unsigned int my_ticket = 0u;                        // thread 0, when this kernel advances the clock
  const int event = clk + 1;
  if (tid == 0) {
    mbar_init(&mbar_tma, 1);
    mbar_init(&mbar_mma, 1);
    mbar_expect_tx(&mbar_tma, (uint32_t)p.tc_blob_bytes + (uint32_t)hdr * 8u);
    tma_bulk_g2s(sT, p.tc_blob, (uint32_t)p.tc_blob_bytes, &mbar_tma);
    tma_bulk_g2s(drow, p.dtab + (size_t)event * p.dstride, (uint32_t)hdr * 8u, &mbar_tma);
    if (p.advance_clock) my_ticket = clock_take_ticket(p.ticket, clk);
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
  // this thread's two 16-byte chunks of A' row `row`: Re(i) of its 4 branches, Im(i) of them
  unsigned char* a_re = sA + (size_t)(row >> 3) * TC_SBO + (size_t)(row & 7) * 16 + grp * 128;
  unsigned char* a_im = a_re + 4 * 128;
  const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);   // lane quadrant of this warp
  uint32_t mma_phase = 0;

  // per-branch constants of this thread's slice
  float u0r[TC_BPT], u0i[TC_BPT], vlo2[TC_BPT], vhi2[TC_BPT];
  int model[TC_BPT];
  bool any_model5 = false;
#pragma unroll
  for (int j = 0; j < TC_BPT; ++j) {
    const int k = grp * TC_BPT + j;
    u0r[j] = u0f[k].x; u0i[j] = u0f[k].y;
    model[j] = bmodel[k];
    vlo2[j] = model[j] == 2 ? 1.f : vlo[k] * vlo[k];
    vhi2[j] = model[j] == 2 ? 1.f : vhi[k] * vhi[k];
    any_model5 |= model[j] == 5;
  }
  any_model5 = __syncthreads_or(any_model5 ? 1 : 0) != 0;     // CTA-uniform
  // loop-invariant UMMA descriptors; K step kk advances the start-address field by 256 B >> 4
  const uint64_t adesc0 = umma_smem_desc(a_addr), b1desc0 = umma_smem_desc(b1_addr),
                 b2desc0 = umma_smem_desc(b2_addr);

  const int tiles = (p.E + TC_M - 1) / TC_M;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int e_raw = tile * TC_M + row;
    const bool valid = e_raw < p.E;
    const int e = valid ? e_raw : p.E - 1;

    // ---- nominal power (p.u. on 1 MVA) and warm start of this thread's branches
    float sr[TC_BPT], si[TC_BPT], dr[TC_BPT], di[TC_BPT];
#pragma unroll
    for (int j = 0; j < TC_BPT; ++j) {
      const int k = grp * TC_BPT + j;
      sr[j] = si[j] = dr[j] = di[j] = 0.f;
      if (k < p.nb) {
        const int l = bload[k];
        float kw, kvar;                                // fp32: this solver's working precision
        if (p.load_kw != nullptr) {
          kw = (float)p.load_kw[(size_t)l * p.E + e];
          kvar = (float)p.load_kvar[(size_t)l * p.E + e];
        } else {
          kw = (float)base_kw[l];
          kvar = (float)base_kvar[l];
          if (p.agent_p != nullptr)                    // multiagent_env.py:171-181, opendss.py:128
            for (int a = 0; a < p.A; ++a)
              if (aslot[a] == l) kw += (float)p.agent_p[(size_t)a * p.E + e];
        }
        const float sh = (float)(share[k] * 1e-3);
        sr[j] = kw * sh;
        si[j] = kvar * sh;
        if (p.warm_start) {
          const double2 up = p.u_state[(size_t)k * p.E + e];
          dr[j] = (float)(up.x - (double)u0r[j]);
          di[j] = (float)(up.y - (double)u0i[j]);
        }
      }
    }

    int it = 0;
    bool conv = !valid, conv_ok = true;
    while (true) {
      // ---- currents of my branches at u = u0 + drop, split hi/lo, into my chunks of A'
      float xr[TC_BPT], xi[TC_BPT];
#pragma unroll
      for (int j = 0; j < TC_BPT; ++j) {
        const float ur = u0r[j] + dr[j], ui = u0i[j] + di[j];
        const float m2 = ur * ur + ui * ui;
        // constant PQ inside the band, constant Z outside = 1 / clamp(|u|^2, vlo^2, vhi^2);
        // a constant-Z load is the degenerate band [1, 1]
        float kf = __fdividef(1.f, fminf(fmaxf(m2, vlo2[j]), vhi2[j]));
        if (any_model5 && model[j] == 5) kf = m2 > 0.f ? rsqrtf(m2) : 0.f;
        xr[j] = (sr[j] * ur + si[j] * ui) * kf;       // conj(s) u k
        xi[j] = (sr[j] * ui - si[j] * ur) * kf;
      }
      {
        const float4 fr = make_float4(xr[0], xr[1], xr[2], xr[3]);
        const float4 fi = make_float4(xi[0], xi[1], xi[2], xi[3]);
        const float4 hr = make_float4(tf32_hi(fr.x), tf32_hi(fr.y), tf32_hi(fr.z), tf32_hi(fr.w));
        const float4 hi = make_float4(tf32_hi(fi.x), tf32_hi(fi.y), tf32_hi(fi.z), tf32_hi(fi.w));
        const float4 lr = make_float4(fr.x - hr.x, fr.y - hr.y, fr.z - hr.z, fr.w - hr.w);
        const float4 li = make_float4(fi.x - hi.x, fi.y - hi.y, fi.z - hi.z, fi.w - hi.w);
        *reinterpret_cast<float4*>(a_re) = hr;                       // K  0..31  x_hi
        *reinterpret_cast<float4*>(a_im) = hi;
        *reinterpret_cast<float4*>(a_re + 1024) = lr;                // K 32..63  x_lo
        *reinterpret_cast<float4*>(a_im + 1024) = li;
        *reinterpret_cast<float4*>(a_re + 2048) = hr;                // K 64..95  x_hi
        *reinterpret_cast<float4*>(a_im + 2048) = hi;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // st.shared -> tensor core
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      const int all_done = __syncthreads_and(conv ? 1 : 0);
      if (all_done) break;                                           // A' = currents at final u

      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int kk = 0; kk < kTcK3 / 8; ++kk)
          umma_tf32(tmem, adesc0 + 16u * kk, b1desc0 + 16u * kk, idesc1, kk > 0 ? 1u : 0u);
        umma_commit(&mbar_mma);
      }
      mbar_wait(&mbar_mma, mma_phase);
      mma_phase ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

      float nr[TC_BPT], ni[TC_BPT];
      tmem_ld4(t_lane + grp * TC_BPT, nr);                           // Re(drop) of my branches
      tmem_ld4(t_lane + kTcNb + grp * TC_BPT, ni);                   // Im(drop)
      float dpart = 0.f;
#pragma unroll
      for (int j = 0; j < TC_BPT; ++j)
        dpart = fmaxf(dpart, fmaxf(fabsf(nr[j] - dr[j]), fabsf(ni[j] - di[j])));
      s_dpart[grp][row] = dpart;
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      if (!conv) {                                                   // per-env convergence mask
        const float d = fmaxf(fmaxf(s_dpart[0][row], s_dpart[1][row]),
                              fmaxf(s_dpart[2][row], s_dpart[3][row]));
#pragma unroll
        for (int j = 0; j < TC_BPT; ++j) { dr[j] = nr[j]; di[j] = ni[j]; }
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
        umma_tf32(tmem + 32, adesc0 + 16u * kk, b2desc0 + 16u * kk, idesc2, kk > 0 ? 1u : 0u);
      umma_commit(&mbar_mma);
    }
    mbar_wait(&mbar_mma, mma_phase);
    mma_phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    float vmn = 3.0e38f, vmx = -3.0e38f;
    for (int c = grp; c < p.tc_n2 / 16; c += 4) {                    // 8 nodes per chunk: (Re, Im) pairs
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
    s_vmn[grp][row] = vmn;
    s_vmx[grp][row] = vmx;
    if (valid) {
#pragma unroll
      for (int j = 0; j < TC_BPT; ++j) {
        const int k = grp * TC_BPT + j;
        if (k < p.nb)
          p.u_state[(size_t)k * p.E + e] =
              make_double2((double)u0r[j] + (double)dr[j], (double)u0i[j] + (double)di[j]);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                        // vmag / partial min-max of the other groups visible

    if (valid && grp == 0) {
      p.vmin[e] = (double)fminf(fminf(s_vmn[0][row], s_vmn[1][row]), fminf(s_vmn[2][row], s_vmn[3][row]));
      p.vmax[e] = (double)fmaxf(fmaxf(s_vmx[0][row], s_vmx[1][row]), fmaxf(s_vmx[2][row], s_vmx[3][row]));
      p.iters[e] = conv_ok ? it : -it;
    }
    if (valid) {
      double pen_share = 0.0, viol = 0.0;
      if (p.punit != 0.0) {
        const double v = p.vmag[(size_t)p.penalty_node * p.E + e];
        viol = fmax(0.0, fmax(p.pvlo - v, v - p.pvhi));
        pen_share = (viol * p.punit) / (double)p.A;
      }
      if (grp == 0) p.viol[e] = viol;
      for (int a = grp; a < p.A; a += 4) {
        const int node = anode[a];
        const size_t ae = (size_t)a * p.E + e;
        p.vbus[ae] = node >= 0 ? p.vmag[(size_t)node * p.E + e] : 1.0;
        if (p.reward_hook) {
          const double r = p.rew[ae] - pen_share;
          p.rew[ae] = r;
          p.rew_copy[ae] = r;
          p.ep_ret[ae] += r;
        }
      }
    }
    __syncthreads();                        // s_* and A' are reused by the next tile
  }

  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_COLS)
                 : "memory");
  if (p.advance_clock && tid == 0)
    clock_advance_if_last(my_ticket, p.ticket, p.clock, clk, p.tickets);
}

int tc_grid(const PfParams& p) {
  const int tiles = (p.E + TC_M - 1) / TC_M;
  const int grid = tiles < 148 * 2 ? tiles : 148 * 2;
  return grid < 1 ? 1 : grid;
}

cudaError_t launch_powerflow_tc(const PfParams& p, cudaStream_t s) {
  const int tiles = (p.E + TC_M - 1) / TC_M;
  const int grid = tc_grid(p);
  const size_t smem = (size_t)TC_A_BYTES + (size_t)p.tc_blob_bytes + (size_t)(2 + 2 * p.nl) * 8 + 16;
  cudaError_t err = cudaFuncSetAttribute(pf_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
  if (err != cudaSuccess) return err;
  pf_tc_kernel<<<grid, TC_THREADS, smem, s>>>(p);
  return cudaGetLastError();
}

}  // namespace pgw
