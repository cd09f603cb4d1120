// K6, tensor-core form for feeders with up to 88 load branches (IEEE-13 ... 123-bus class): the
// Z-bus fixed point  u <- u0 - Zbb i(u)  as a dense FP16 contraction on tcgen05 with FP32
// accumulation in TMEM, one CTA per tile of 128 envs, Zbb resident in shared memory.
//
//   D[env, :] = X[env, :] * B^T       M = 128 envs (TMEM lanes),  N = K = 16 * NCH,
//                                      NCH = chunks of 8 branches (instantiated: 2, 4, 8, 11)
//
// * Real-ified complex product, interleaved by chunks of 8 branches: columns 16c..16c+7 hold
//   Re, 16c+8..16c+15 hold Im of branches 8c..8c+7 - on the K side (currents) as well as on
//   the N side (voltage drops) - so one K = 16 MMA step is one chunk, and one 16-column
//   tcgen05.ld hands a thread the complex drops of the 8 branches it owns.
// * Split FP16: x = x_hi + x_lo, B = B_hi + B_lo (each part 11 significant bits, both operands
//   pre-scaled by powers of two into the FP16 range).  The three significant products
//   x_lo B_hi + x_hi B_lo + x_hi B_hi are ONE accumulation chain of 3 * NCH MMAs that re-uses
//   the two resident images of each operand (no tripled K).  Error of a drop: ~1e-8 p.u.
// * Shared memory (NCH = 11): B_hi | B_lo 121 kB, A_hi | A_lo 88 kB, tables 8 kB.  All images
//   are canonical K-major no-swizzle UMMA tiles (8-row x 16-byte core matrices, LBO 128 B,
//   SBO 256 * NCH B); B images are prepared on the host and staged by TMA bulk copies, A is
//   written by the epilogue threads (16-byte st.shared + fence.proxy.async).  Per-branch
//   constants come through the constant bank (a __grid_constant__ parameter), because the MMA
//   operand reads alone need ~110 of the 128 B/clk of shared-memory bandwidth.
// * TMEM: two accumulators D[0], D[1] of N columns used alternately, so the previous drop is
//   still there for the per-env convergence test max|du| < tol and nothing but the nominal
//   powers lives in registers across iterations.  Converged envs stop rewriting their row of
//   A (convergence mask): their drop reproduces itself.
// * Iteration = pass 1 (max|du| over my chunks, CTA barrier, mask) + pass 2 (currents at the new
//   voltages, chunk by chunk).  As soon as the G chunks of a wave are written by all warps
//   (mbarrier) the issuer warp feeds those K steps of the NEXT chain into the other
//   accumulator: the tensor core runs while the remaining chunks are evaluated.
// * Expansion v = w - Znb i: nodes that ARE a load-branch voltage up to a real factor (wye
//   loads) come from u directly; the rest go through the same chain against row chunks of Znb,
//   streamed from L2 over the B images by TMA (next chunk prefetched during the epilogue) or
//   kept resident for small feeders; then |v|, min/max, agent bus voltages, the reward hook.
//
// Same inputs/outputs as pf_fixed_point_kernel (powerflow.cu).
#include <cuda_fp16.h>

#include <type_traits>

#include "internal.cuh"
#include "tc2_common.cuh"
#include "tma.cuh"

namespace pgw {

// Phase stamps of tools/phase_probe.py (instrumented build only, -DPGW_PHASE_TIMERS): thread 0 of
// every CTA records the SM clock at the phase boundaries of its FIRST tile.
#ifdef PGW_PHASE_TIMERS
#define T2_STAMP(k)                                                          \
  do {                                                                       \
    if (p.phase_clk != nullptr && threadIdx.x == 0 && stamp_tile)            \
      p.phase_clk[(size_t)blockIdx.x * 16 + (k)] = clock64();                \
  } while (0)
#else
#define T2_STAMP(k) do { } while (0)
#endif

// POLISH: after the tensor-core fixed point has converged, `t.polish` sweeps of the same fixed
// point in float64 on the SIMT pipe (Zbb, u0 and the load tables in float64 from the FP64
// solver's feeder blob, nominal powers kept in shared memory since the prologue, currents
// exchanged through shared memory).  The tensor core has taken the solution to ~5e-8 p.u. (the
// split-FP16 operands carry 22 bits); every float64 sweep multiplies that error by the
// contraction factor of the iteration (~0.2 on IEEE-13), so two sweeps reach ~3e-9 and three
// ~6e-10 -- what the shared-penalty hook (1e4 x violation) needs for rewards within 2e-5.
template <int NCH, bool ANY_M5, bool STANDALONE, int OCC = t2_ctas_per_sm(NCH), bool POLISH = false>
__global__ void __launch_bounds__(NCH <= 4 ? 256 : 512, OCC)
    pf_tc2_kernel(const PfParams p, const __grid_constant__ Tc2Consts kc,
                  const __grid_constant__ std::conditional_t<POLISH, Tc2Polish, int> kp) {
  static_assert(!POLISH || (NCH == 2 && !STANDALONE), "FP64 polish: <= 16 load branches, step solve");
  constexpr int G = NCH <= 4 ? 2 : 4;                  // groups of 128 threads
  constexpr int T2_THREADS = 128 * G;
  constexpr int T2_ISSUER = 4 * G - 1;                 // warp that feeds the pipelined chains
  constexpr int SLOTS = (NCH + G - 1) / G;             // chunks per thread: c = grp + G * slot
  constexpr int N = 16 * NCH;
  constexpr uint32_t SBO = 256u * NCH;                 // bytes between 8-row groups
  constexpr uint32_t PB = (uint32_t)(N / 8) * SBO;     // one B image: N rows
  constexpr uint32_t APB = 16u * SBO;                  // one A image: 128 rows
  extern __shared__ __align__(1024) unsigned char t2_smem[];
  __shared__ __align__(8) uint64_t mbar_tab, mbar_b, mbar_mma, mbar_wave[SLOTS];
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_dpart[2][G][T2_M];                // per-group partial max |d drop|, 2 phases
  float (*s_vmn)[T2_M] = s_dpart[0], (*s_vmx)[T2_M] = s_dpart[1];   // reused after the solve
  const Tc2Params& t = p.tc2;
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);            // provably warp-uniform
  const int row = tid & (T2_M - 1);                    // env row in the tile = TMEM lane
  const int grp = warp >> 2;                           // chunk group 0..G-1
  const int hdr = 2 + 2 * p.nl;

  unsigned char* sB = t2_smem;                         // B_hi | B_lo (iteration) or a Znb chunk
  // small feeders: every Znb chunk stays resident behind the Zbb images (no streaming, all
  // expansion chains issued back to back into their own accumulators)
  const int nres = t.resident ? t.ncc : 0;
  unsigned char* sA = sB + (size_t)(1 + nres) * 2 * PB;   // A_hi | A_lo
  unsigned char* sT = sA + 2 * APB;                    // tables
  double* drow = reinterpret_cast<double*>(sT + t.tab_bytes);
  // POLISH: nominal powers s64 [16][128] (thread private, kept since the prologue) and the
  // currents of a sweep i64 [16][128], both float64 complex
  double2* sS64 = reinterpret_cast<double2*>(drow + hdr);
  double2* sI64 = sS64 + (size_t)16 * T2_M;

  [[maybe_unused]] bool stamp_tile = true;
  [[maybe_unused]] int it_stamp = 0;
#ifdef PGW_PHASE_TIMERS
  if (p.phase_clk != nullptr && threadIdx.x == 0) p.phase_clk[(size_t)blockIdx.x * 16 + 12] = (long long)global_timer_ns();
#endif
  T2_STAMP(0);
  const int clk = p.event_mode == 0 ? -1 : *p.clock;
  unsigned int my_ticket = 0u;                        // thread 0, when this kernel advances the clock
  const int event = clk + 1;
  T2_STAMP(1);
  if (tid == 0) {
    mbar_init(&mbar_tab, 1);
    mbar_init(&mbar_b, 1);
    mbar_init(&mbar_mma, 1);
    for (int w = 0; w < SLOTS; ++w)                    // one arrival per warp that owns a chunk of wave w
      mbar_init(&mbar_wave[w], 4u * (uint32_t)(NCH - G * w < G ? NCH - G * w : G));
    mbar_expect_tx(&mbar_tab, (uint32_t)t.tab_bytes + (uint32_t)hdr * 8u);
    tma_bulk_g2s(sT, t.blob + t.off_tab, (uint32_t)t.tab_bytes, &mbar_tab);
    mbar_expect_tx(&mbar_b, (uint32_t)(1 + nres) * 2 * PB);
    tma_bulk_g2s(sB, t.blob, (uint32_t)(1 + nres) * 2 * PB, &mbar_b);   // chunks follow in the blob
    // last: the event row's address waits for the device clock
    tma_bulk_g2s(drow, p.dtab + (size_t)event * p.dstride, (uint32_t)hdr * 8u, &mbar_tab);
    if (p.advance_clock) my_ticket = clock_take_ticket(p.ticket, clk);
  }
  // Pull the tile's per-env inputs (warm-start voltages, agent powers) towards L2 while the
  // tables, the operand images and the TMEM allocation are in flight: 32-byte sectors of
  // [rows][128 envs], spread over the CTA.
  auto prefetch_tile = [&](int tile) {
    const size_t e0 = (size_t)p.e_lo + (size_t)tile * T2_M;
    if (p.warm_start)
      for (int i = tid; i < p.nb * 64; i += T2_THREADS) {
        const size_t ee = e0 + (size_t)(i & 63) * 2;
        if (ee < (size_t)p.e_hi)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(p.u_state + (size_t)(i >> 6) * p.E + ee));
      }
    if (!STANDALONE && p.agent_p != nullptr)
      for (int i = tid; i < p.A * 32; i += T2_THREADS) {
        const size_t ee = e0 + (size_t)(i & 31) * 4;
        if (ee < (size_t)p.e_hi) {
          const size_t off = (size_t)(i >> 5) * p.E + ee;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(p.agent_p + off));
          if (p.reward_hook) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.ep_ret + off));
        }
      }
  };
  prefetch_tile(blockIdx.x);
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
  T2_STAMP(2);
  mbar_wait(&mbar_tab, 0);
  T2_STAMP(3);
  // Programmatic dependent launch: everything above overlapped the tail of the component kernel;
  // its outputs (agent powers, rewards) are visible from here on.
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // per-branch constants come through the constant bank (kernel parameter), not shared memory:
  // the tensor core's operand reads already take most of the shared-memory bandwidth
  const float* share = reinterpret_cast<const float*>(sT + t.t_share);
  const int32_t* bload = reinterpret_cast<const int32_t*>(sT + t.t_bload);
  const int32_t* bagent = reinterpret_cast<const int32_t*>(sT + t.t_bagent);
  const float2* wx = reinterpret_cast<const float2*>(sT + t.t_w);        // w of the expanded slots
  const int32_t* xnode = reinterpret_cast<const int32_t*>(sT + t.t_xnode);  // slot -> node or -1
  const int32_t* dnode = reinterpret_cast<const int32_t*>(sT + t.t_dnode);  // branch -> node or -1
  const float* dscale = reinterpret_cast<const float*>(sT + t.t_dscale);
  const int2* vag = reinterpret_cast<const int2*>(sT + t.t_vag);   // branch -> up to two agents whose
  const int32_t* vtail = reinterpret_cast<const int32_t*>(sT + t.t_vtail);  // bus node it derives; rest
  const int32_t* lptr = reinterpret_cast<const int32_t*>(sT + t.t_lptr);
  const int32_t* lidx = reinterpret_cast<const int32_t*>(sT + t.t_lidx);
  const int32_t* anode = reinterpret_cast<const int32_t*>(sT + t.t_anode);
  const double* base_kw = drow + 2;
  const double* base_kvar = drow + 2 + p.nl;

  const uint32_t idesc = t2_idesc_f16(N);
  const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
  // this thread's 16-byte Re slot of chunk 0 in A_hi; chunk c is +256 c, Im +128, A_lo +APB
  unsigned char* a_row = sA + (size_t)(row >> 3) * SBO + (size_t)(row & 7) * 16;
  const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);   // lane quadrant of this warp
  const float xs = t.xscale, ds1 = t.descale1, ds2 = t.descale2;
  const float tol_s = t.tol / ds1;                     // tolerance in accumulator units
  uint32_t mma_phase = 0, wave_phase = 0, b_phase = 0; // b_phase is used by warp 0 only
  bool b_fresh = true;                                 // a load of the Zbb images is in flight

  // One accumulation chain D = A_lo B_hi + A_hi B_lo + A_hi B_hi (small terms first), issued by
  // one elected lane of warp 0; all operands are warp-uniform (uniform registers in SASS).
  auto issue_chain = [&](uint32_t d_col, bool wait_b, uint32_t b_off = 0u, bool commit = true) {
    if (warp == 0) {
      if (t2_elect_one()) {
        if (wait_b) mbar_wait(&mbar_b, b_phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t a_hi = t2_smem_desc(sA_u, SBO), a_lo = t2_smem_desc(sA_u + APB, SBO);
        const uint64_t b_hi = t2_smem_desc(sB_u + b_off, SBO),
                       b_lo = t2_smem_desc(sB_u + b_off + PB, SBO);
#pragma unroll
        for (int kk = 0; kk < NCH; ++kk)
          t2_umma_f16(tmem + d_col, a_lo + 16u * kk, b_hi + 16u * kk, idesc, kk > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < NCH; ++kk)
          t2_umma_f16(tmem + d_col, a_hi + 16u * kk, b_lo + 16u * kk, idesc, 1u);
#pragma unroll
        for (int kk = 0; kk < NCH; ++kk)
          t2_umma_f16(tmem + d_col, a_hi + 16u * kk, b_hi + 16u * kk, idesc, 1u);
        if (commit) t2_commit(&mbar_mma);
      }
      if (wait_b) b_phase ^= 1u;
      __syncwarp();
    }
  };
  auto load_b = [&](const unsigned char* src) {        // stage two images over sB (thread 0)
    mbar_expect_tx(&mbar_b, 2 * PB);
    tma_bulk_g2s(sB, src, 2 * PB, &mbar_b);
  };

  const int tiles = (p.e_hi - p.e_lo + T2_M - 1) / T2_M;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int e_raw = p.e_lo + tile * T2_M + row;
    const bool valid = e_raw < p.e_hi;
    const int e = valid ? e_raw : p.e_hi - 1;
    const bool more_tiles = tile + (int)gridDim.x < tiles;

    // ---- nominal power (p.u. on 1 MVA, scaled by xs) of my branches; initial drop into D[1].
    //      All global loads of a chunk are issued before any of them is used.
    float sr[SLOTS][8], si[SLOTS][8];
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      const int c = grp + G * s;
      if (c < NCH) {
        float d0[16], x[8], y[8];
#pragma unroll
        for (int h = 0; h < 8; h += 4) {               // half chunks: 12 loads in flight per thread
          double kwd[4], kvd[4];
          double2 up[4];
          int ld[4], ag[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            ld[j] = bload[8 * c + h + j];
            ag[j] = bagent[8 * c + h + j];
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = 8 * c + h + j;
            kwd[j] = 0.0;
            kvd[j] = 0.0;
            if (STANDALONE) {
              if (k < p.nb) {
                kwd[j] = p.load_kw[(size_t)ld[j] * p.E + e];
                kvd[j] = p.load_kvar[(size_t)ld[j] * p.E + e];
              }
            } else if (p.agent_p != nullptr && ag[j] >= 0) {
              kwd[j] = p.agent_p[(size_t)ag[j] * p.E + e];
            }
            const float4 c0 = t2_cst(kc, k);
            up[j] = make_double2((double)c0.x, (double)c0.y);
            if (p.warm_start && k < p.nb) up[j] = p.u_state[(size_t)k * p.E + e];
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = 8 * c + h + j, jj = h + j;
            if (!STANDALONE) {
              double kw = base_kw[ld[j]];
              // multiagent_env.py:171-181: P summed per load name in agent order, then added to
              // the scaled base load (opendss.py:128); ag == -2: several agents on this load
              if (p.agent_p != nullptr && ag[j] == -2)
                kw += t2_shared_load_kw(p.agent_p, lidx, lptr[ld[j]], lptr[ld[j] + 1], p.E, e);
              kwd[j] += kw;
              kvd[j] = base_kvar[ld[j]];
            }
            const float sh = share[k] * xs;            // 0 for the padded branches
            sr[s][jj] = (float)kwd[j] * sh;
            si[s][jj] = (float)kvd[j] * sh;
            if constexpr (POLISH)                      // nominal power in float64 for the polish
              sS64[(size_t)k * T2_M + row] = make_double2(kwd[j] * kp.share[k], kvd[j] * kp.share[k]);
            const float4 cc = t2_cst(kc, k);
            d0[jj] = ((float)up[j].x - cc.x) * (1.f / ds1);        // initial guess: fp32 is plenty
            d0[8 + jj] = ((float)up[j].y - cc.y) * (1.f / ds1);
            t2_current<ANY_M5>(cc, ANY_M5 ? t2_gh(kc, k) : make_float2(1.f, 0.f), d0[jj], d0[8 + jj], ds1,
                               sr[s][jj], si[s][jj], x[jj], y[jj]);
          }
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

    // ---- fixed point.  Per iteration:
    //   wait for the chain of this iteration (drop n in D[cur], drop n-1 still in D[cur^1])
    //   pass 1: max |d drop| over my chunks -> CTA barrier -> per-env convergence mask
    //   pass 2: chunk by chunk, currents at the new voltages into A; as soon as the four chunks
    //           of a wave are written by all warps, the issuer warp feeds the K steps of those
    //           chunks of the NEXT chain into D[cur^1] (free since the barrier), so the tensor
    //           core works while the remaining chunks are still being evaluated.
    T2_STAMP(4);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // st.shared -> tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    issue_chain(0u, b_fresh);                          // first chain; the Zbb images have landed
    T2_STAMP(5);
    int it = 0, my_it = 0, cur = 0;                    // D[cur] = newest drop once its chain is done
    bool conv = !valid, conv_ok = true;
    while (true) {
      mbar_wait(&mbar_mma, mma_phase);
      mma_phase ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      ++it;

      float dpart = 0.f;
#pragma unroll
      for (int s = 0; s < SLOTS; ++s) {
        const int c = grp + G * s;
        if (c < NCH) {                                 // warp-uniform: tcgen05.ld is collective
          float dn[16], dold[16];
          t2_ld16(t_lane + cur * N + 16 * c, dn);
          t2_ld16(t_lane + (cur ^ 1) * N + 16 * c, dold);
#pragma unroll
          for (int j = 0; j < 16; j += 2)
            dpart = fmaxf(dpart, fmaxf(fabsf(dn[j] - dold[j]), fabsf(dn[j + 1] - dold[j + 1])));
        }
      }
      s_dpart[it & 1][grp][row] = dpart;
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      const bool all_done =
          __syncthreads_and((conv || dpart < tol_s || it >= p.max_iter) ? 1 : 0) != 0;
      const bool frozen = conv;                        // converged before this iteration
      if (!conv) {                                     // per-env convergence mask
        float d = s_dpart[it & 1][0][row];
#pragma unroll
        for (int q = 1; q < G; ++q) d = fmaxf(d, s_dpart[it & 1][q][row]);
        conv_ok = d < tol_s;
        conv = conv_ok || it >= p.max_iter;
        my_it = it;
      }

#pragma unroll
      for (int s = 0; s < SLOTS; ++s) {
        const int c = grp + G * s;
        if (c < NCH) {
          float dn[16], x[8], y[8];
          t2_ld16(t_lane + cur * N + 16 * c, dn);
          if (!frozen) {                               // the env's last write holds x(u_final)
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {           // branch pairs 8c + 2jp, 8c + 2jp + 1
              const int q = 4 * c + jp;
              float2 x2, y2;
              t2_current2<ANY_M5>(kc.pa[q], kc.pb[q], ANY_M5 ? kc.pc[q] : make_float4(1.f, 1.f, 0.f, 0.f),
                                  make_float2(dn[2 * jp], dn[2 * jp + 1]),
                                  make_float2(dn[8 + 2 * jp], dn[9 + 2 * jp]), ds1,
                                  make_float2(sr[s][2 * jp], sr[s][2 * jp + 1]),
                                  make_float2(si[s][2 * jp], si[s][2 * jp + 1]), x2, y2);
              x[2 * jp] = x2.x; x[2 * jp + 1] = x2.y;
              y[2 * jp] = y2.x; y[2 * jp + 1] = y2.y;
            }
            uint4 hi, lo;
            t2_split8(x, hi, lo);
            *reinterpret_cast<uint4*>(a_row + 256 * c) = hi;
            *reinterpret_cast<uint4*>(a_row + 256 * c + APB) = lo;
            t2_split8(y, hi, lo);
            *reinterpret_cast<uint4*>(a_row + 256 * c + 128) = hi;
            *reinterpret_cast<uint4*>(a_row + 256 * c + 128 + APB) = lo;
          }
          if (!all_done) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&mbar_wave[s]);
          }
        }
        if (!all_done && warp == T2_ISSUER) {          // K steps of wave s of the next chain
          if (t2_elect_one()) {
            mbar_wait(&mbar_wave[s], wave_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d_col = (uint32_t)((cur ^ 1) * N);
            const uint64_t a_hi = t2_smem_desc(sA_u, SBO), a_lo = t2_smem_desc(sA_u + APB, SBO);
            const uint64_t b_hi = t2_smem_desc(sB_u, SBO), b_lo = t2_smem_desc(sB_u + PB, SBO);
#pragma unroll
            for (int kk = G * s; kk < G * s + G && kk < NCH; ++kk) {
              t2_umma_f16(tmem + d_col, a_lo + 16u * kk, b_hi + 16u * kk, idesc, kk > 0 ? 1u : 0u);
              t2_umma_f16(tmem + d_col, a_hi + 16u * kk, b_lo + 16u * kk, idesc, 1u);
              t2_umma_f16(tmem + d_col, a_hi + 16u * kk, b_hi + 16u * kk, idesc, 1u);
            }
            if (s == SLOTS - 1) t2_commit(&mbar_mma);
          }
          __syncwarp();
        }
      }
      if (all_done) break;                             // A = currents at the final voltages
      wave_phase ^= 1u;
      cur ^= 1;
    }
    T2_STAMP(6);
    if (stamp_tile) it_stamp = it;
    const int last = cur;                              // D[last] = final drop
    cur ^= 1;                                          // the expansion starts in the older buffer

    // ---- expansion to all node voltages, Znb row chunks streamed over the B images
    if (tid == 0 && t.ncc > 0 && !t.resident) load_b(t.blob + t.off_zn);
    b_fresh = t.ncc > 0 && !t.resident;                // the Zbb images are restaged after the chunks
    if (more_tiles) prefetch_tile(tile + (int)gridDim.x);
    // Final branch voltages: warm-start state, and the magnitudes of the nodes that ARE a
    // load branch voltage up to a real factor (wye loads: v_node = dscale * u_branch).
    float vmn = 3.0e38f, vmx = -3.0e38f;
    [[maybe_unused]] bool chains_issued = false;
    const bool polish = POLISH && t.polish > 0;
    if (polish) {
      // (a) The currents at the final voltages are in A: the expansion chains of a small feeder
      //     (every Znb chunk resident, accumulators of their own) start now and run on the tensor
      //     core while the SIMT pipe polishes the branch voltages.
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      if (t.resident && t.ncc > 0) {
        for (int cc = 0; cc < t.ncc; ++cc)
          issue_chain((uint32_t)((2 + cc) * N), false, (uint32_t)(1 + cc) * 2 * PB, cc + 1 == t.ncc);
        chains_issued = true;
      }
      // (b) float64 sweeps u <- u0 - Zbb i(u) over my chunk's branches (G == 2, one chunk per
      //     thread): t.polish full sweeps, then one over the rows the rewards read.  The env's
      //     currents go through shared memory, [branch][env row] (conflict-free 16-byte accesses).
      if constexpr (POLISH) {
        const int c = grp;
        double2 u[8];
        {
          // Sweep 0 in float32 (as in step_fused.cu): the full sweep feeds the currents of the row
          // sweep below, which contracts its error by > 10x; float32 currents halve the
          // shared-memory traffic and the 14 x 8 complex MACs per thread run on the FP32 pipe
          // instead of the FP64 one (262 144 IEEE-13 envs: the float64 form of this sweep was most
          // of the 78 us the polish added to the 104 us solve).
          float dn[16];
          t2_ld16(t_lane + last * N + 16 * c, dn);
          float2* sI32 = reinterpret_cast<float2*>(sI64);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k = 8 * c + j;
            float x, y;
            t2_current32<ANY_M5>(t2_cst(kc, k), ANY_M5 ? t2_gh(kc, k) : make_float2(1.f, 0.f), dn[j], dn[8 + j],
                                 ds1, sr[0][j], si[0][j], x, y);
            if (k < p.nb) sI32[(size_t)k * T2_M + row] = make_float2(x, y);
          }
          __syncthreads();
          T2_STAMP(14);
          float2 acc[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll 2
          for (int jj = 0; jj < p.nb; ++jj) {
            const float2 ij = sI32[(size_t)jj * T2_M + row];
            const float2* z = kp.z32 + jj * 16 + 8 * c;
#pragma unroll
            for (int j = 0; j < 8; ++j) t2_cmac_sub32(acc[j], z[j], ij);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const double2 b = kp.u0[8 * c + j];
            u[j] = make_double2(b.x + (double)acc[j].x, b.y + (double)acc[j].y);
            if (t.polish == 1 && valid && 8 * c + j < p.nb) p.u_state[(size_t)(8 * c + j) * p.E + e] = u[j];
          }
          T2_STAMP(15);
        }
        const int sweeps = t.polish + ((kp.rows != 0u || t.polish_row >= 0) ? 1 : 0);
#pragma unroll 1
        for (int sweep = 1; sweep < sweeps; ++sweep) {
          __syncthreads();                             // everyone has read the previous currents
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k = 8 * c + j;
            if (k < p.nb)
              sI64[(size_t)k * T2_M + row] = t2_current64(kp.model[k], sS64[(size_t)k * T2_M + row], u[j],
                                                          kp.vlo2[k], kp.vhi2[k]);
          }
          __syncthreads();
          if (sweep < t.polish) {
            t2_sweep64(kp, p.nb, c, sI64, row, u);
            if (sweep + 1 == t.polish) {               // warm-start state: after the full sweeps
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (valid && 8 * c + j < p.nb) p.u_state[(size_t)(8 * c + j) * p.E + e] = u[j];
            }
          } else if ((kp.rows >> (8 * c)) & 0xffu) {   // warp uniform; rolled loops (cold, few rows)
#pragma unroll 1
            for (int j = 0; j < 8; ++j) {
              const int k = 8 * c + j;
              if (!((kp.rows >> k) & 1u)) continue;
              double2 acc = kp.u0[k];
#pragma unroll 2
              for (int jj = 0; jj < p.nb; ++jj)
                t2_cmac_sub64(acc, kp.zT[jj * 16 + k], sI64[(size_t)jj * T2_M + row]);
#pragma unroll
              for (int q = 0; q < 8; ++q)              // u stays in registers: no dynamic index
                if (q == j) u[q] = acc;
            }
          }
        }
        // (c) the wye-load nodes from the polished voltages
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = 8 * c + j;
          const int n = dnode[k];
          if (n >= 0) {
            const double m2 = u[j].x * u[j].x + u[j].y * u[j].y;
            double mag;
            if ((kp.rows >> k) & 1u) {
              mag = sqrt(m2) * kp.dscale[k];
            } else {                                   // float32 rsqrt + one Newton step: ~1e-14
              const double r = (double)t2_rsqrt((float)m2);
              const double y = m2 * r;
              mag = fma(fma(-y, y, m2), 0.5 * r, y) * kp.dscale[k];
            }
            vmn = fminf(vmn, (float)mag);
            vmx = fmaxf(vmx, (float)mag);
            if (valid) {
              p.vmag[(size_t)n * p.E + e] = mag;
              if (!p.reward_hook) {
                const int2 ag2 = vag[k];
                if (ag2.x >= 0) p.vbus[(size_t)ag2.x * p.E + e] = mag;
                if (ag2.y >= 0) p.vbus[(size_t)ag2.y * p.E + e] = mag;
              }
            }
          }
        }
      }
    } else {
#pragma unroll 1                                       // executed once per tile: keep the code small
      for (int s = 0; s < SLOTS; ++s) {
        const int c = grp + G * s;
        if (c < NCH) {                                   // warp-uniform: tcgen05.ld is collective
          float dn[16];
          t2_ld16(t_lane + last * N + 16 * c, dn);
  #pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k = 8 * c + j;
            const float4 c0 = t2_cst(kc, k);
            const float ur = fmaf(dn[j], ds1, c0.x), ui = fmaf(dn[8 + j], ds1, c0.y);
            if (valid && k < p.nb)
              p.u_state[(size_t)k * p.E + e] = make_double2((double)ur, (double)ui);
            const int n = dnode[k];
            if (n >= 0) {
              const float m2 = fmaf(ur, ur, ui * ui);
              const float mag = m2 * t2_rsqrt(fmaxf(m2, 1e-30f)) * dscale[k];
              vmn = fminf(vmn, mag);
              vmx = fmaxf(vmx, mag);
              if (valid) {
                p.vmag[(size_t)n * p.E + e] = (double)mag;
                if (!p.reward_hook) {                    // bus voltage of the agents at this node
                  const int2 ag2 = vag[k];
                  if (ag2.x >= 0) p.vbus[(size_t)ag2.x * p.E + e] = (double)mag;
                  if (ag2.y >= 0) p.vbus[(size_t)ag2.y * p.E + e] = (double)mag;
                }
              }
            }
          }
        }
      }
    }
    T2_STAMP(7);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                                   // D[last] may be overwritten from chunk 1 on

    if (t.resident && t.ncc > 0) {                     // accumulators 2 + cc, one commit for all
      if (!chains_issued)
        for (int cc = 0; cc < t.ncc; ++cc)
          issue_chain((uint32_t)((2 + cc) * N), false, (uint32_t)(1 + cc) * 2 * PB, cc + 1 == t.ncc);
      mbar_wait(&mbar_mma, mma_phase);
      mma_phase ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    for (int cc = 0; cc < t.ncc; ++cc) {
      const int dsel = t.resident ? 2 + cc : (cur + cc) & 1;
      if (!t.resident) {
        issue_chain((uint32_t)(dsel * N), true);
        mbar_wait(&mbar_mma, mma_phase);
        mma_phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0 && (cc + 1 < t.ncc || more_tiles))  // the B images are free again
          load_b(cc + 1 < t.ncc ? t.blob + t.off_zn + (size_t)(cc + 1) * 2 * PB : t.blob);
      }
#pragma unroll 1
      for (int s = 0; s < SLOTS; ++s) {
        const int c = grp + G * s;
        if (c < NCH) {
          const int s0 = 8 * (cc * NCH + c);           // expanded slots s0 .. s0 + 7
          if (s0 < t.nx) {
            float v[16];
            t2_ld16(t_lane + dsel * N + 16 * c, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int n = xnode[s0 + j];
              if (n >= 0) {
                const float2 w = wx[s0 + j];
                const float vr = fmaf(v[j], ds2, w.x), vi = fmaf(v[8 + j], ds2, w.y);
                const float m2 = fmaf(vr, vr, vi * vi);
                const float mag = m2 * t2_rsqrt(fmaxf(m2, 1e-30f));
                vmn = fminf(vmn, mag);
                vmx = fmaxf(vmx, mag);
                if (valid) p.vmag[(size_t)n * p.E + e] = (double)mag;
              }
            }
          }
        }
      }
      if (!t.resident) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                               // this accumulator is rewritten two chunks on
      }
    }
    s_vmn[grp][row] = vmn;
    s_vmx[grp][row] = vmx;
    T2_STAMP(8);
    __syncthreads();                        // vmag / partial min-max of the other groups visible

    if (valid && grp == 0) {
      float mn = s_vmn[0][row], mx = s_vmx[0][row];
#pragma unroll
      for (int q = 1; q < G; ++q) { mn = fminf(mn, s_vmn[q][row]); mx = fmaxf(mx, s_vmx[q][row]); }
      p.vmin[e] = (double)mn;
      p.vmax[e] = (double)mx;
      p.iters[e] = conv_ok ? my_it : -my_it;
    }
    // Penalty node that is not a wye-load node (delta load, or no load at all): its row of the
    // last sweep, v = w - Znb[row] i(u), from the currents still in shared memory.
    [[maybe_unused]] double v_row = 0.0;
    const bool have_row = POLISH && polish && t.polish_row >= 0;
    if (POLISH && have_row) {
      const double2* wv = reinterpret_cast<const double2*>(p.blob + p.off_w);
      const double2* zn = reinterpret_cast<const double2*>(p.blob + p.off_znbT);
      double2 v = wv[t.polish_row];
      for (int jj = 0; jj < p.nb; ++jj)
        t2_cmac_sub64(v, zn[(size_t)jj * p.nnp + t.polish_row], sI64[(size_t)jj * T2_M + row]);
      v_row = sqrt(v.x * v.x + v.y * v.y);
      if (valid && grp == 0) p.vmag[(size_t)t.polish_row * p.E + e] = v_row;
    }
    if (valid) {
      double pen_share = 0.0, viol = 0.0;
      if (p.punit != 0.0) {
        const double v = have_row ? v_row : p.vmag[(size_t)p.penalty_node * p.E + e];
        viol = fmax(0.0, fmax(p.pvlo - v, v - p.pvhi));
        pen_share = (viol * p.punit) / (double)p.A;
      }
      if (grp == 0) p.viol[e] = viol;
      if (p.reward_hook) {
        // every agent: bus voltage + the shared penalty.  Agents a = grp, grp + G, ... in
        // batches of 4 so that the loads of a batch are in flight together.
        for (int a0 = grp; a0 < p.A; a0 += 4 * G) {
          double vb[4], rw[4], er[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int a = a0 + G * q;
            vb[q] = 1.0; rw[q] = 0.0; er[q] = 0.0;
            if (a < p.A) {
              const int node = anode[a];
              const size_t ae = (size_t)a * p.E + e;
              if (node >= 0) vb[q] = (have_row && node == t.polish_row) ? v_row : p.vmag[(size_t)node * p.E + e];
              rw[q] = p.rew[ae];
              er[q] = p.ep_ret[ae];
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int a = a0 + G * q;
            if (a < p.A) {
              const size_t ae = (size_t)a * p.E + e;
              const double r = rw[q] - pen_share;
              p.vbus[ae] = vb[q];
              p.rew[ae] = r;
              p.rew_copy[ae] = r;
              p.ep_ret[ae] = er[q] + r;
            }
          }
        }
      } else {
        // the agents whose bus is not a wye-load node (none in the shipped scenarios)
        for (int q = grp; q < t.ntail; q += G) {
          const int a = vtail[q], node = anode[a];
          p.vbus[(size_t)a * p.E + e] = node >= 0 ? p.vmag[(size_t)node * p.E + e] : 1.0;
        }
      }
    }
    __syncthreads();                        // s_* and A are reused by the next tile
    T2_STAMP(9);
    stamp_tile = false;
  }

  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem),
                 "r"((uint32_t)t.tmem_cols)
                 : "memory");
  if (p.advance_clock && tid == 0)
    clock_advance_if_last(my_ticket, p.ticket, p.clock, clk, p.tickets);
#ifdef PGW_PHASE_TIMERS
  if (p.phase_clk != nullptr && threadIdx.x == 0) {
    p.phase_clk[(size_t)blockIdx.x * 16 + 10] = clock64();
    p.phase_clk[(size_t)blockIdx.x * 16 + 11] = (long long)it_stamp;
    p.phase_clk[(size_t)blockIdx.x * 16 + 13] = (long long)global_timer_ns();
  }
#endif
}

size_t tc2_smem_bytes(const PfParams& p) {
  return (size_t)(1 + (p.tc2.resident ? p.tc2.ncc : 0)) * 2 * p.tc2.part_bytes +
         (size_t)2 * 16 * 256 * p.tc2.nch + (size_t)p.tc2.tab_bytes +
         (size_t)(2 + 2 * p.nl) * 8 + 16 + tc2_polish_bytes(p);
}

// FP64 polish area: s64 and i64, [16][128] complex each
size_t tc2_polish_bytes(const PfParams& p) {
  if (!tc2_polish_active(p)) return 0;
  return (size_t)2 * 16 * T2_M * sizeof(double2);
}

// the polish runs in the step solve of a feeder with <= 16 load branches whose rewards depend on
// the fresh voltages (shared-penalty hook)
bool tc2_polish_active(const PfParams& p) {
  return p.tc2.polish > 0 && p.tc2.nch == 2 && p.reward_hook && p.load_kw == nullptr && p.event_mode == 1;
}

int tc2_padded_chunks(int nch) {            // instantiated tile widths
  return nch <= 2 ? 2 : nch <= 4 ? 4 : nch <= 8 ? 8 : nch <= 11 ? 11 : 0;
}

template <bool M5>
static cudaError_t launch_tc2_polish(const PfParams& p, int grid, size_t smem, cudaStream_t s) {
  auto kern = pf_tc2_kernel<2, M5, false, 2, true>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = p.pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, p, *p.tc2.consts, *p.tc2.pconsts);
}

template <int NCH, int OCC>
static cudaError_t launch_tc2_t(const PfParams& p, int grid, size_t smem, cudaStream_t s) {
  auto kern = p.load_kw != nullptr
                  ? (p.tc2.any_m5 ? pf_tc2_kernel<NCH, true, true, OCC> : pf_tc2_kernel<NCH, false, true, OCC>)
                  : (p.tc2.any_m5 ? pf_tc2_kernel<NCH, true, false, OCC> : pf_tc2_kernel<NCH, false, false, OCC>);
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NCH <= 4 ? 256 : 512);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = p.pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, p, *p.tc2.consts, 0);
}

int tc2_grid(const PfParams& p) {
  const int tiles = (p.e_hi - p.e_lo + T2_M - 1) / T2_M;
  const bool polish = tc2_polish_active(p);
  const bool dense = !polish && p.tc2.nch == 2 && tiles > 148 * t2_ctas_per_sm(2) && p.tc2.tmem_cols <= 128;
  const int per_sm = dense ? 4 : t2_ctas_per_sm(p.tc2.nch);
  const int grid = tiles < 148 * per_sm ? tiles : 148 * per_sm;
  return grid < 1 ? 1 : grid;
}

cudaError_t launch_powerflow_tc2(const PfParams& p, cudaStream_t s) {
  const int tiles = (p.e_hi - p.e_lo + T2_M - 1) / T2_M;
  // the dense build pays off when the tiles do not fit one wave of the regular one
  const bool polish = tc2_polish_active(p);
  const bool dense = !polish && p.tc2.nch == 2 && tiles > 148 * t2_ctas_per_sm(2) && p.tc2.tmem_cols <= 128;
  const int per_sm = dense ? 4 : t2_ctas_per_sm(p.tc2.nch);
  int grid = tiles < 148 * per_sm ? tiles : 148 * per_sm;
  if (grid < 1) grid = 1;
  const size_t smem = tc2_smem_bytes(p);
  if (polish) return p.tc2.any_m5 ? launch_tc2_polish<true>(p, grid, smem, s) : launch_tc2_polish<false>(p, grid, smem, s);
  switch (p.tc2.nch) {
    case 2: return dense ? launch_tc2_t<2, 4>(p, grid, smem, s) : launch_tc2_t<2, 2>(p, grid, smem, s);
    case 4: return launch_tc2_t<4, 2>(p, grid, smem, s);
    case 8: return launch_tc2_t<8, 1>(p, grid, smem, s);
    case 11: return launch_tc2_t<11, 1>(p, grid, smem, s);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace pgw
