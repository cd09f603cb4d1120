// K6 (FP64 SIMT form) + the fresh-voltage half of K7.
//
// Batched three-phase Z-bus fixed point over the load branches of a compiled feeder:
//     u <- u0 - Zbb * i(u),      i_k(u_k) = load characteristic of branch k
// followed by the expansion to all node voltages  v = w - Znb * i,  per-unit
// magnitudes, per-env min/max, the voltage at each agent's bus node, and the
// grid-level reward hook (shared voltage-violation penalty).
//
// Replaces `Solve mode=snap` + `_prepare_bus_voltages`
// (gridworld/distribution_system/opendss.py:134, :156-165), the per-load kW/kvar
// bookkeeping (:107-131), `get_external_obs_vars` (gridworld/multiagent_env.py:90-115)
// and `CoordinatedMultiBuildingControlEnv.reward_transform`
// (examples/marl/openai/train.py:51-88).
//
// Mapping: TPE threads cooperate on one env (branch rows strided over the lanes, R rows
// per lane), so a 13-bus env (14 branches) is half a warp and convergence masking is a
// per-half-warp predicate.  The feeder tables (Zbb, Znb, u0, w, branch table; 13.6 kB for
// IEEE-13) and the event row are staged into shared memory by TMA bulk copies once per
// CTA; for the half-warp mapping each lane additionally keeps its row of Zbb in
// registers.  Branch currents are exchanged through shared memory.  The solve warm-starts
// from the env's previous solution.  Bound by FP64 FMA latency, not by HBM.
#include "internal.cuh"
#include "tma.cuh"

namespace pgw {

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// Current drawn by one load branch at branch voltage u (per unit), nominal power s:
// OpenDSS load model 1 = constant PQ inside [vmin, vmax], constant Z outside
// (continuous at the band edges); 2 = constant Z; 5 = constant current magnitude.
__device__ __forceinline__ double2 branch_current(int model, double2 s, double2 u, double vmin,
                                                  double vmax) {
  const double m2 = u.x * u.x + u.y * u.y;
  const double2 sc = make_double2(s.x, -s.y);            // conj(s) = yeq on a 1 p.u. base
  double k;
  if (model == 2) {
    k = 1.0;
  } else if (model == 5) {
    k = m2 > 0.0 ? rsqrt(m2) : 0.0;
  } else if (m2 <= vmin * vmin) {
    k = __drcp_rn(vmin * vmin);
  } else if (m2 > vmax * vmax) {
    k = __drcp_rn(vmax * vmax);
  } else {
    k = __drcp_rn(m2);                                   // conj(s / u) = conj(s) u / |u|^2
  }
  const double2 t = cmul(sc, u);
  return make_double2(t.x * k, t.y * k);
}

template <int TPE>
__device__ __forceinline__ double group_max(double v, unsigned mask) {
#pragma unroll
  for (int o = TPE / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(mask, v, o, TPE));
  return v;
}
template <int TPE>
__device__ __forceinline__ double group_min(double v, unsigned mask) {
#pragma unroll
  for (int o = TPE / 2; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(mask, v, o, TPE));
  return v;
}

// acc -= z * i  (complex), as four FMAs
__device__ __forceinline__ void cmac_sub(double2& acc, double2 z, double2 i) {
  acc.x = fma(-z.x, i.x, acc.x);
  acc.x = fma(z.y, i.y, acc.x);
  acc.y = fma(-z.x, i.y, acc.y);
  acc.y = fma(-z.y, i.x, acc.y);
}

// 2 CTAs of 256 threads per SM (<= 128 registers): 4096 envs = 256 CTAs fit in one wave.
template <int TPE, int R, bool ZREG>
__global__ void __launch_bounds__(256, 2) pf_fixed_point_kernel(const PfParams p) {
  extern __shared__ __align__(16) unsigned char pf_smem[];
  __shared__ __align__(8) uint64_t mbar;
  const int EPB = blockDim.x / TPE;                      // envs per block pass
  const int lane = threadIdx.x % TPE;
  const int le = threadIdx.x / TPE;
  const int nbp = p.nbp;
  const int hdr = 2 + 2 * p.nl;                          // done, reserved, kW[nl], kvar[nl]

  // shared memory: [feeder blob (if staged)] [event row header] [currents] [|v| staging]
  const int blob_smem = p.stage_blob ? p.blob_bytes : 0;
  double* drow = reinterpret_cast<double*>(pf_smem + blob_smem);
  double2* icur_all = reinterpret_cast<double2*>(pf_smem + blob_smem + (size_t)hdr * 8);
  double* stage_all = reinterpret_cast<double*>(icur_all + (size_t)EPB * nbp);
  double2* icur = icur_all + (size_t)le * nbp;
  double* stage = stage_all + (size_t)le * p.nn;
  const unsigned gmask =
      TPE == 32 ? 0xffffffffu : (((1u << TPE) - 1u) << (TPE * ((threadIdx.x & 31) / TPE)));

  const int clk = p.event_mode == 0 ? -1 : *p.clock;
  unsigned int my_ticket = 0u;                        // thread 0, when this kernel advances the clock
  const int event = clk + 1;
  if (threadIdx.x == 0) {
    mbar_init(&mbar, 1);
    mbar_expect_tx(&mbar, (uint32_t)blob_smem + (uint32_t)hdr * 8u);
    if (p.stage_blob) tma_bulk_g2s(pf_smem, p.blob, (uint32_t)p.blob_bytes, &mbar);
    tma_bulk_g2s(drow, p.dtab + (size_t)event * p.dstride, (uint32_t)hdr * 8u, &mbar);
    if (p.advance_clock) my_ticket = clock_take_ticket(p.ticket, clk);
  }
  __syncthreads();
  mbar_wait(&mbar, 0);

  const unsigned char* tab = p.stage_blob ? pf_smem : p.blob;
  const double2* zbbT = reinterpret_cast<const double2*>(tab);
  const double2* u0t = reinterpret_cast<const double2*>(tab + p.off_u0);
  const double2* znbT = reinterpret_cast<const double2*>(tab + p.off_znbT);
  const double2* wt = reinterpret_cast<const double2*>(tab + p.off_w);
  const double* share = reinterpret_cast<const double*>(tab + p.off_share);
  const double* vminpu = reinterpret_cast<const double*>(tab + p.off_vmin);
  const double* vmaxpu = reinterpret_cast<const double*>(tab + p.off_vmax);
  const int32_t* bload = reinterpret_cast<const int32_t*>(tab + p.off_bload);
  const int32_t* bmodel = reinterpret_cast<const int32_t*>(tab + p.off_bmodel);
  const int32_t* aslot = reinterpret_cast<const int32_t*>(tab + p.off_slot);
  const int32_t* anode = reinterpret_cast<const int32_t*>(tab + p.off_node);
  const double* base_kw = drow + 2;
  const double* base_kvar = drow + 2 + p.nl;

  // this lane's row of Zbb in registers (half-warp / one-row-per-lane mapping only)
  double2 zr[ZREG ? TPE : 1];
  if (ZREG) {
#pragma unroll
    for (int j = 0; j < TPE; ++j) zr[j] = zbbT[(size_t)j * nbp + lane];
  }

  const int groups = (p.e_hi - p.e_lo + EPB - 1) / EPB;
  for (int g = blockIdx.x; g < groups; g += gridDim.x) {
    const int e_raw = p.e_lo + g * EPB + le;
    const bool valid = e_raw < p.e_hi;
    const int e = valid ? e_raw : p.e_hi - 1;

    // ---- per-branch nominal power: base load of the event + the agents on that load
    double2 s[R], u[R], u0[R];
    double vlo[R], vhi[R];
    int model[R];
#pragma unroll
    for (int m = 0; m < R; ++m) {
      const int k = lane + TPE * m;
      s[m] = make_double2(0.0, 0.0);
      u0[m] = u0t[k];
      u[m] = u0[m];
      model[m] = 1; vlo[m] = 0.95; vhi[m] = 1.05;
      if (k < p.nb) {
        const int l = bload[k];
        double kw, kvar;
        if (p.load_kw != nullptr) {
          kw = p.load_kw[(size_t)l * p.E + e];
          kvar = p.load_kvar[(size_t)l * p.E + e];
        } else {
          kw = base_kw[l];
          kvar = base_kvar[l];
        }
        if (p.load_kw == nullptr && p.agent_p != nullptr) {
          // multiagent_env.py:171-181: P summed per load name in agent order, then added
          // to the scaled base load (opendss.py:128)
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
        const double sh = share[k] * 1e-3;               // kVA -> p.u. on 1 MVA
        s[m] = make_double2(kw * sh, kvar * sh);
        model[m] = bmodel[k];
        vlo[m] = vminpu[k];
        vhi[m] = vmaxpu[k];
        if (p.warm_start) u[m] = p.u_state[(size_t)k * p.E + e];
      }
    }

    // ---- fixed point with per-env convergence masking
    int it = 0;
    bool conv = false;
    while (true) {
#pragma unroll
      for (int m = 0; m < R; ++m)
        icur[lane + TPE * m] = branch_current(model[m], s[m], u[m], vlo[m], vhi[m]);
      __syncwarp(gmask);
      double2 acc[R], acc2[R];
#pragma unroll
      for (int m = 0; m < R; ++m) { acc[m] = u0[m]; acc2[m] = make_double2(0.0, 0.0); }
      if (ZREG) {
#pragma unroll
        for (int j = 0; j < TPE; j += 2) {               // padded columns of Zbb are zero
          cmac_sub(acc[0], zr[j], icur[j]);
          cmac_sub(acc2[0], zr[j + 1], icur[j + 1]);
        }
      } else {
        int j = 0;
        for (; j + 1 < p.nb; j += 2) {
          const double2 i0 = icur[j], i1 = icur[j + 1];
          const double2* z0 = zbbT + (size_t)j * nbp + lane;
          const double2* z1 = z0 + nbp;
#pragma unroll
          for (int m = 0; m < R; ++m) {
            cmac_sub(acc[m], z0[TPE * m], i0);
            cmac_sub(acc2[m], z1[TPE * m], i1);
          }
        }
        if (j < p.nb) {
          const double2 i0 = icur[j];
          const double2* z0 = zbbT + (size_t)j * nbp + lane;
#pragma unroll
          for (int m = 0; m < R; ++m) cmac_sub(acc[m], z0[TPE * m], i0);
        }
      }
      double d = 0.0;
#pragma unroll
      for (int m = 0; m < R; ++m) {
        acc[m].x += acc2[m].x;
        acc[m].y += acc2[m].y;
        d = fmax(d, fmax(fabs(acc[m].x - u[m].x), fabs(acc[m].y - u[m].y)));
        u[m] = acc[m];
      }
      d = group_max<TPE>(d, gmask);
      ++it;
      conv = d < p.tol;
      if (conv || it >= p.max_iter) break;
      __syncwarp(gmask);
    }
    if (valid) {
#pragma unroll
      for (int m = 0; m < R; ++m) {
        const int k = lane + TPE * m;
        if (k < p.nb) p.u_state[(size_t)k * p.E + e] = u[m];
      }
    }

    // ---- all node voltages from the currents of the last iteration
    double lmin = 1e300, lmax = -1e300;
    for (int n = lane; n < p.nn; n += TPE) {
      double2 v = wt[n], v2 = make_double2(0.0, 0.0);
      int k = 0;
      for (; k + 1 < p.nb; k += 2) {
        cmac_sub(v, znbT[(size_t)k * p.nnp + n], icur[k]);
        cmac_sub(v2, znbT[(size_t)(k + 1) * p.nnp + n], icur[k + 1]);
      }
      if (k < p.nb) cmac_sub(v, znbT[(size_t)k * p.nnp + n], icur[k]);
      v.x += v2.x;
      v.y += v2.y;
      const double mag = sqrt(v.x * v.x + v.y * v.y);
      stage[n] = mag;
      lmin = fmin(lmin, mag);
      lmax = fmax(lmax, mag);
    }
    lmin = group_min<TPE>(lmin, gmask);
    lmax = group_max<TPE>(lmax, gmask);
    __syncwarp(gmask);

    if (valid) {
      if (lane == 0) {
        p.vmin[e] = lmin;
        p.vmax[e] = lmax;
        p.iters[e] = conv ? it : -it;
      }
      double pen_share = 0.0;
      if (p.punit != 0.0) {
        const double v = stage[p.penalty_node];
        const double viol = fmax(0.0, fmax(p.pvlo - v, v - p.pvhi));   // train.py:71-88
        if (lane == 0) p.viol[e] = viol;
        pen_share = (viol * p.punit) / (double)p.A;                    // train.py:56-61
      } else if (lane == 0) {
        p.viol[e] = 0.0;
      }
      for (int a = lane; a < p.A; a += TPE) {
        const int node = anode[a];
        const size_t ae = (size_t)a * p.E + e;
        p.vbus[ae] = node >= 0 ? stage[node] : 1.0;
        if (p.reward_hook) {
          const double r = p.rew[ae] - pen_share;
          p.rew[ae] = r;
          p.rew_copy[ae] = r;
          p.ep_ret[ae] += r;
        }
      }
    }
    // ---- coalesced store of the magnitudes: [nn][E], EPB consecutive envs per row
    __syncthreads();
    for (int idx = threadIdx.x; idx < p.nn * EPB; idx += blockDim.x) {
      const int n = idx / EPB, j = idx % EPB;
      const int ee = p.e_lo + g * EPB + j;
      if (ee < p.e_hi) p.vmag[(size_t)n * p.E + ee] = stage_all[(size_t)j * p.nn + n];
    }
    __syncthreads();
  }
  if (p.advance_clock && threadIdx.x == 0)
    clock_advance_if_last(my_ticket, p.ticket, p.clock, clk, p.tickets);
}

int fp64_grid(const PfParams& p);

template <int TPE, int R, bool ZREG>
static cudaError_t launch_pf_t(const PfParams& p, cudaStream_t s) {
  const int threads = 256;
  const int epb = threads / TPE;
  const int grid = fp64_grid(p);
  const size_t smem = (size_t)(p.stage_blob ? p.blob_bytes : 0) + (size_t)(2 + 2 * p.nl) * 8 +
                      (size_t)epb * p.nbp * sizeof(double2) + (size_t)epb * p.nn * sizeof(double);
  if (smem > 48 * 1024) {
    cudaError_t err = cudaFuncSetAttribute(pf_fixed_point_kernel<TPE, R, ZREG>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
  }
  pf_fixed_point_kernel<TPE, R, ZREG><<<grid, threads, smem, s>>>(p);
  return cudaGetLastError();
}

int fp64_grid(const PfParams& p) {
  const int tpe = p.nbp == 16 ? 16 : 32;
  const int epb = 256 / tpe;
  const int groups = (p.e_hi - p.e_lo + epb - 1) / epb;
  const int grid = groups < 148 * 4 ? groups : 148 * 4;
  return grid < 1 ? 1 : grid;
}

cudaError_t launch_powerflow(const PfParams& p, cudaStream_t s) {
  if (p.nbp == 16) return launch_pf_t<16, 1, true>(p, s);
  if (p.nbp == 32) return launch_pf_t<32, 1, false>(p, s);
  if (p.nbp == 64) return launch_pf_t<32, 2, false>(p, s);
  if (p.nbp == 96) return launch_pf_t<32, 3, false>(p, s);
  if (p.nbp == 128) return launch_pf_t<32, 4, false>(p, s);
  return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------ K8: episode statistics
__global__ void __launch_bounds__(256) stats_kernel(const StatsParams p) {
  __shared__ double red[5][8];
  __shared__ double mn[8], mx[8];
  double rsum = 0.0, esum = 0.0, vsum = 0.0, nconv = 0.0, itsum = 0.0;
  double vmn = 1e300, vmx = -1e300;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < p.E; e += gridDim.x * blockDim.x) {
    for (int a = 0; a < p.A; ++a) {
      rsum += p.rew[(size_t)a * p.E + e];
      esum += p.ep_ret[(size_t)a * p.E + e];
    }
    if (p.iters != nullptr) {
      vsum += p.viol[e];
      const int it = p.iters[e];
      nconv += it < 0 ? 1.0 : 0.0;
      itsum += (double)(it < 0 ? -it : it);
      vmn = fmin(vmn, p.vmin[e]);
      vmx = fmax(vmx, p.vmax[e]);
    }
  }
  double vals[5] = {rsum, esum, vsum, nconv, itsum};
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int q = 0; q < 5; ++q) vals[q] += __shfl_xor_sync(0xffffffffu, vals[q], o);
    vmn = fmin(vmn, __shfl_xor_sync(0xffffffffu, vmn, o));
    vmx = fmax(vmx, __shfl_xor_sync(0xffffffffu, vmx, o));
  }
  const int warp = threadIdx.x >> 5, ln = threadIdx.x & 31;
  if (ln == 0) {
    for (int q = 0; q < 5; ++q) red[q][warp] = vals[q];
    mn[warp] = vmn;
    mx[warp] = vmx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[5] = {0, 0, 0, 0, 0}, a = 1e300, b = -1e300;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      for (int q = 0; q < 5; ++q) t[q] += red[q][w];
      a = fmin(a, mn[w]);
      b = fmax(b, mx[w]);
    }
    atomicAdd(p.out + 1, t[0]);
    atomicAdd(p.out + 2, t[1]);
    atomicAdd(p.out + 3, t[2]);
    atomicAdd(p.out + 4, t[3]);
    atomicAdd(p.out + 5, t[4]);
    // min / max through the ordered-integer view of positive doubles
    atomicMin(reinterpret_cast<unsigned long long*>(p.out + 6),
              (unsigned long long)__double_as_longlong(a));
    atomicMax(reinterpret_cast<unsigned long long*>(p.out + 7),
              (unsigned long long)__double_as_longlong(b > 0 ? b : 0.0));
    if (blockIdx.x == 0) p.out[0] = (double)p.E * (double)(*p.clock);
  }
}

cudaError_t launch_stats(const StatsParams& p, cudaStream_t s) {
  int grid = (p.E + 255) / 256;
  if (grid > 148 * 4) grid = 148 * 4;
  stats_kernel<<<grid, 256, 0, s>>>(p);
  return cudaGetLastError();
}

}  // namespace pgw
