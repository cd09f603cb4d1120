// K6, tensor-core form for 123-bus-class feeders (up to 88 load branches): the Z-bus fixed
// point  u <- u0 - Zbb i(u)  as a dense FP16 contraction on tcgen05 with FP32 accumulation in
// TMEM, one CTA per tile of 128 envs, Zbb resident in shared memory for the whole solve.
//
//   D[env, :] = X[env, :] * B^T       M = 128 envs (TMEM lanes),  N = K = 16 * nch,
//                                      nch = ceil(nb / 8) chunks of 8 branches
//
// * Real-ified complex product, interleaved by chunks of 8 branches: columns 16c..16c+7 hold
//   Re, 16c+8..16c+15 hold Im of branches 8c..8c+7 - on the K side (currents) as well as on
//   the N side (voltage drops) - so one K = 16 MMA step is one chunk, and one 16-column
//   tcgen05.ld hands a thread the complex drops of the 8 branches it owns.
// * Split FP16: x = x_hi + x_lo, B = B_hi + B_lo (each part 11 significant bits, both operands
//   pre-scaled by powers of two into the FP16 range).  The three significant products
//   x_lo B_hi + x_hi B_lo + x_hi B_hi are ONE accumulation chain of 3 * nch MMAs that re-uses
//   the two resident images of each operand (no tripled K).  Error of a drop: ~1e-8 p.u.
// * Shared memory (nch = 11): B_hi | B_lo 121 kB, A_hi | A_lo 88 kB, tables 8 kB.  All images
//   are canonical K-major no-swizzle UMMA tiles (8-row x 16-byte core matrices, LBO 128 B,
//   SBO 256 * nch B); B images are prepared on the host and staged by TMA bulk copies, A is
//   written by the epilogue threads (16-byte st.shared + fence.proxy.async).
// * TMEM: two accumulators D[0], D[1] of N columns used alternately, so the previous drop is
//   still there for the per-env convergence test max|du| < tol and nothing but the nominal
//   powers lives in registers across iterations.  Converged envs stop rewriting their row of
//   A (convergence mask): their drop reproduces itself.
// * Expansion v = w - Znb i for all nodes: the same chain against ceil(2 nn / N) row chunks
//   of Znb streamed from L2 over the B images by TMA (prefetch of chunk c+1 overlaps the
//   epilogue of chunk c), then |v|, min/max, agent bus voltages and the reward hook.
//
// Same inputs/outputs as pf_fixed_point_kernel (powerflow.cu).
#include <cuda_fp16.h>

#include "internal.cuh"
#include "tma.cuh"

namespace pgw {

constexpr int T2_M = 128;
constexpr int T2_THREADS = 512;
constexpr int T2_SLOTS = 3;                 // chunks per thread: c = grp + 4 * slot

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

__global__ void __launch_bounds__(T2_THREADS, 1) pf_tc2_kernel(const PfParams p) {
  extern __shared__ __align__(1024) unsigned char t2_smem[];
  __shared__ __align__(8) uint64_t mbar_tab, mbar_b, mbar_mma;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_dpart[2][4][T2_M];                // per-group partial max |d drop|, 2 phases
  float (*s_vmn)[T2_M] = s_dpart[0], (*s_vmx)[T2_M] = s_dpart[1];   // reused after the solve
  const Tc2Params& t = p.tc2;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & (T2_M - 1);                    // env row in the tile = TMEM lane
  const int grp = tid >> 7;                            // chunk group 0..3
  const int nch = t.nch;
  const int N = 16 * nch;
  const uint32_t sbo = 256u * (uint32_t)nch;           // bytes between 8-row groups
  const uint32_t PB = (uint32_t)t.part_bytes;          // one B image: N rows
  const uint32_t APB = 16u * sbo;                      // one A image: 128 rows
  const int hdr = 2 + 2 * p.nl;

  unsigned char* sB = t2_smem;                         // B_hi | B_lo (iteration) or a Znb chunk
  unsigned char* sA = sB + 2 * PB;                     // A_hi | A_lo
  unsigned char* sT = sA + 2 * APB;                    // tables
  double* drow = reinterpret_cast<double*>(sT + t.tab_bytes);

  const int event = p.event_mode == 0 ? 0 : (*p.clock + 1);
  if (tid == 0) {
    mbar_init(&mbar_tab, 1);
    mbar_init(&mbar_b, 1);
    mbar_init(&mbar_mma, 1);
    mbar_expect_tx(&mbar_tab, (uint32_t)t.tab_bytes + (uint32_t)hdr * 8u);
    tma_bulk_g2s(sT, t.blob + t.off_tab, (uint32_t)t.tab_bytes, &mbar_tab);
    tma_bulk_g2s(drow, p.dtab + (size_t)event * p.dstride, (uint32_t)hdr * 8u, &mbar_tab);
    mbar_expect_tx(&mbar_b, 2 * PB);
    tma_bulk_g2s(sB, t.blob, 2 * PB, &mbar_b);
  }
  if (warp == 0) {                                     // one warp owns TMEM alloc / free
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_s)),
                 "r"((uint32_t)t.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  mbar_wait(&mbar_tab, 0);

  const float2* u0f = reinterpret_cast<const float2*>(sT + t.t_u0);
  const float* vlo2 = reinterpret_cast<const float*>(sT + t.t_vlo2);
  const float* vhi2 = reinterpret_cast<const float*>(sT + t.t_vhi2);
  const float* share = reinterpret_cast<const float*>(sT + t.t_share);
  const int32_t* m5 = reinterpret_cast<const int32_t*>(sT + t.t_m5);
  const int32_t* bload = reinterpret_cast<const int32_t*>(sT + t.t_bload);
  const float2* wf = reinterpret_cast<const float2*>(sT + t.t_w);
  const int32_t* lptr = reinterpret_cast<const int32_t*>(sT + t.t_lptr);
  const int32_t* lidx = reinterpret_cast<const int32_t*>(sT + t.t_lidx);
  const int32_t* anode = reinterpret_cast<const int32_t*>(sT + t.t_anode);
  const double* base_kw = drow + 2;
  const double* base_kvar = drow + 2 + p.nl;

  const uint32_t idesc = t2_idesc_f16(N);
  const uint64_t a_hi = t2_smem_desc(smem_u32(sA), sbo), a_lo = t2_smem_desc(smem_u32(sA + APB), sbo);
  const uint64_t b_hi = t2_smem_desc(smem_u32(sB), sbo), b_lo = t2_smem_desc(smem_u32(sB + PB), sbo);
  // this thread's 16-byte Re slot of chunk 0 in A_hi; chunk c is +256 c, Im +128, A_lo +APB
  unsigned char* a_row = sA + (size_t)(row >> 3) * sbo + (size_t)(row & 7) * 16;
  const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);   // lane quadrant of this warp
  const float xs = t.xscale, ds1 = t.descale1, ds2 = t.descale2;
  const float tol_s = t.tol / ds1;                     // tolerance in accumulator units
  uint32_t mma_phase = 0, b_phase = 0;                 // b_phase is tracked by thread 0 only

  // One accumulation chain D = A_lo B_hi + A_hi B_lo + A_hi B_hi (small terms first).
  auto issue_chain = [&](uint32_t d_col) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int kk = 0; kk < nch; ++kk)
      t2_umma_f16(tmem + d_col, a_lo + 16u * kk, b_hi + 16u * kk, idesc, kk > 0 ? 1u : 0u);
    for (int kk = 0; kk < nch; ++kk)
      t2_umma_f16(tmem + d_col, a_hi + 16u * kk, b_lo + 16u * kk, idesc, 1u);
    for (int kk = 0; kk < nch; ++kk)
      t2_umma_f16(tmem + d_col, a_hi + 16u * kk, b_hi + 16u * kk, idesc, 1u);
    t2_commit(&mbar_mma);
  };

  const int tiles = (p.E + T2_M - 1) / T2_M;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int e_raw = tile * T2_M + row;
    const bool valid = e_raw < p.E;
    const int e = valid ? e_raw : p.E - 1;
    const bool more_tiles = tile + (int)gridDim.x < tiles;

    // ---- nominal power (p.u. on 1 MVA, scaled by xs) of my branches; initial drop into D[1]
    float sr[T2_SLOTS][8], si[T2_SLOTS][8];
#pragma unroll
    for (int s = 0; s < T2_SLOTS; ++s) {
      const int c = grp + 4 * s;
      if (c < nch) {
        float d0[16], x[8], y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = 8 * c + j;
          sr[s][j] = si[s][j] = 0.f;
          d0[j] = d0[8 + j] = 0.f;
          if (k < p.nb) {
            const int l = bload[k];
            float kw, kvar;                            // fp32: this solver's working precision
            if (p.load_kw != nullptr) {
              kw = (float)p.load_kw[(size_t)l * p.E + e];
              kvar = (float)p.load_kvar[(size_t)l * p.E + e];
            } else {
              double kwd = base_kw[l];
              kvar = (float)base_kvar[l];
              if (p.agent_p != nullptr)                // multiagent_env.py:171-181, opendss.py:128
                for (int q = lptr[l]; q < lptr[l + 1]; ++q)
                  kwd += p.agent_p[(size_t)lidx[q] * p.E + e];
              kw = (float)kwd;
            }
            const float sh = share[k] * xs;
            sr[s][j] = kw * sh;
            si[s][j] = kvar * sh;
            if (p.warm_start) {
              const double2 up = p.u_state[(size_t)k * p.E + e];
              d0[j] = (float)(up.x - (double)u0f[k].x);
              d0[8 + j] = (float)(up.y - (double)u0f[k].y);
            }
          }
          const float ur = u0f[k].x + d0[j], ui = u0f[k].y + d0[8 + j];
          const float m2 = ur * ur + ui * ui;
          float kf = __fdividef(1.f, fminf(fmaxf(m2, vlo2[k]), vhi2[k]));
          if (t.any_m5 && m5[k]) kf = m2 > 0.f ? rsqrtf(m2) : 0.f;
          x[j] = (sr[s][j] * ur + si[s][j] * ui) * kf;
          y[j] = (sr[s][j] * ui - si[s][j] * ur) * kf;
          d0[j] *= 1.f / ds1;
          d0[8 + j] *= 1.f / ds1;
        }
        uint4 hi, lo;
        t2_split8(x, hi, lo);
        *reinterpret_cast<uint4*>(a_row + 256 * c) = hi;
        *reinterpret_cast<uint4*>(a_row + 256 * c + APB) = lo;
        t2_split8(y, hi, lo);
        *reinterpret_cast<uint4*>(a_row + 256 * c + 128) = hi;
        *reinterpret_cast<uint4*>(a_row + 256 * c + 128 + APB) = lo;
        t2_st16(t_lane + N + 16 * c, d0);
      }
    }

    int it = 0, my_it = 0, cur = 0;                    // D[cur] receives the next drop
    bool conv = !valid, conv_ok = true;
    float dpart = 3.0e38f;
    while (true) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // st.shared -> tensor core
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      s_dpart[it & 1][grp][row] = dpart;
      const int all_done = __syncthreads_and((conv || dpart < tol_s || it >= p.max_iter) ? 1 : 0);
      if (it > 0 && !conv) {                           // per-env convergence mask
        const float d = fmaxf(fmaxf(s_dpart[it & 1][0][row], s_dpart[it & 1][1][row]),
                              fmaxf(s_dpart[it & 1][2][row], s_dpart[it & 1][3][row]));
        conv_ok = d < tol_s;
        conv = conv_ok || it >= p.max_iter;
        my_it = it;
      }
      if (all_done) break;                             // A = currents at the final u

      if (tid == 0) {
        if (it == 0) { mbar_wait(&mbar_b, b_phase); b_phase ^= 1u; }     // Zbb images landed
        issue_chain((uint32_t)(cur * N));
      }
      mbar_wait(&mbar_mma, mma_phase);
      mma_phase ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

      dpart = 0.f;
#pragma unroll
      for (int s = 0; s < T2_SLOTS; ++s) {
        const int c = grp + 4 * s;
        if (c < nch) {
          float dn[16], x[8], y[8];
          {
            float dold[16];
            t2_ld16(t_lane + cur * N + 16 * c, dn);
            t2_ld16(t_lane + (cur ^ 1) * N + 16 * c, dold);
#pragma unroll
            for (int j = 0; j < 16; ++j) dpart = fmaxf(dpart, fabsf(dn[j] - dold[j]));
          }
          if (!conv) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int k = 8 * c + j;
              const float ur = fmaf(dn[j], ds1, u0f[k].x), ui = fmaf(dn[8 + j], ds1, u0f[k].y);
              const float m2 = ur * ur + ui * ui;
              // constant PQ inside the band, constant Z outside = 1 / clamp(|u|^2, vlo^2, vhi^2);
              // a constant-Z load is the degenerate band [1, 1]
              float kf = __fdividef(1.f, fminf(fmaxf(m2, vlo2[k]), vhi2[k]));
              if (t.any_m5 && m5[k]) kf = m2 > 0.f ? rsqrtf(m2) : 0.f;
              x[j] = (sr[s][j] * ur + si[s][j] * ui) * kf;       // conj(s) u k
              y[j] = (sr[s][j] * ui - si[s][j] * ur) * kf;
            }
            uint4 hi, lo;
            t2_split8(x, hi, lo);
            *reinterpret_cast<uint4*>(a_row + 256 * c) = hi;
            *reinterpret_cast<uint4*>(a_row + 256 * c + APB) = lo;
            t2_split8(y, hi, lo);
            *reinterpret_cast<uint4*>(a_row + 256 * c + 128) = hi;
            *reinterpret_cast<uint4*>(a_row + 256 * c + 128 + APB) = lo;
          }
        }
      }
      ++it;
      cur ^= 1;
    }
    const int last = cur ^ 1;                          // D[last] = final drop

    // ---- expansion to all node voltages, Znb row chunks streamed over the B images
    if (tid == 0) {
      mbar_expect_tx(&mbar_b, 2 * PB);
      tma_bulk_g2s(sB, t.blob + t.off_zn, 2 * PB, &mbar_b);
    }
#pragma unroll
    for (int s = 0; s < T2_SLOTS; ++s) {
      const int c = grp + 4 * s;
      if (c < nch) {                                   // warp-uniform: tcgen05.ld is collective
        float dn[16];
        t2_ld16(t_lane + last * N + 16 * c, dn);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = 8 * c + j;
          if (valid && k < p.nb)
            p.u_state[(size_t)k * p.E + e] =
                make_double2((double)u0f[k].x + (double)(dn[j] * ds1),
                             (double)u0f[k].y + (double)(dn[8 + j] * ds1));
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                                   // D[last] may be overwritten from chunk 1 on

    float vmn = 3.0e38f, vmx = -3.0e38f;
    for (int cc = 0; cc < t.ncc; ++cc) {
      const int dsel = (cur + cc) & 1;
      if (tid == 0) {
        mbar_wait(&mbar_b, b_phase);
        b_phase ^= 1u;
        issue_chain((uint32_t)(dsel * N));
      }
      mbar_wait(&mbar_mma, mma_phase);
      mma_phase ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (tid == 0 && (cc + 1 < t.ncc || more_tiles)) {               // the B images are free again
        mbar_expect_tx(&mbar_b, 2 * PB);
        tma_bulk_g2s(sB, cc + 1 < t.ncc ? t.blob + t.off_zn + (size_t)(cc + 1) * 2 * PB : t.blob,
                     2 * PB, &mbar_b);
      }
#pragma unroll
      for (int s = 0; s < T2_SLOTS; ++s) {
        const int c = grp + 4 * s;
        if (c < nch) {
          const int n0 = 8 * (cc * nch + c);
          if (n0 < p.nn) {
            float v[16];
            t2_ld16(t_lane + dsel * N + 16 * c, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int n = n0 + j;
              if (n < p.nn) {
                const float vr = fmaf(v[j], ds2, wf[n].x), vi = fmaf(v[8 + j], ds2, wf[n].y);
                const float mag = __fsqrt_rn(vr * vr + vi * vi);
                vmn = fminf(vmn, mag);
                vmx = fmaxf(vmx, mag);
                if (valid) p.vmag[(size_t)n * p.E + e] = (double)mag;
              }
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();                                 // this accumulator is rewritten two chunks on
    }
    s_vmn[grp][row] = vmn;
    s_vmx[grp][row] = vmx;
    __syncthreads();                        // vmag / partial min-max of the other groups visible

    if (valid && grp == 0) {
      p.vmin[e] = (double)fminf(fminf(s_vmn[0][row], s_vmn[1][row]), fminf(s_vmn[2][row], s_vmn[3][row]));
      p.vmax[e] = (double)fmaxf(fmaxf(s_vmx[0][row], s_vmx[1][row]), fmaxf(s_vmx[2][row], s_vmx[3][row]));
      p.iters[e] = conv_ok ? my_it : -my_it;
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
        if (p.event_mode != 0) {
          const double r = p.rew[ae] - pen_share;
          p.rew[ae] = r;
          p.rew_copy[ae] = r;
          p.ep_ret[ae] += r;
        }
      }
    }
    __syncthreads();                        // s_* and A are reused by the next tile
  }

  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem),
                 "r"((uint32_t)t.tmem_cols)
                 : "memory");
  if (p.advance_clock) publish_clock_last_cta(p.ticket, p.clock, event, gridDim.x);
}

size_t tc2_smem_bytes(const PfParams& p) {
  return (size_t)2 * p.tc2.part_bytes + (size_t)2 * 16 * 256 * p.tc2.nch + (size_t)p.tc2.tab_bytes +
         (size_t)(2 + 2 * p.nl) * 8 + 16;
}

cudaError_t launch_powerflow_tc2(const PfParams& p, cudaStream_t s) {
  const int tiles = (p.E + T2_M - 1) / T2_M;
  int grid = tiles < 148 ? tiles : 148;
  if (grid < 1) grid = 1;
  const size_t smem = tc2_smem_bytes(p);
  cudaError_t err = cudaFuncSetAttribute(pf_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
  if (err != cudaSuccess) return err;
  pf_tc2_kernel<<<grid, T2_THREADS, smem, s>>>(p);
  return cudaGetLastError();
}

}  // namespace pgw
