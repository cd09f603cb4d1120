// Per-(env, agent) arithmetic of the component models, shared by the step and reset
// kernels of components.cu.  Everything here is float64 and is compiled with
// -fmad=false so that each operation rounds exactly like the NumPy/Python scalar
// arithmetic of the reference (the storage clamp, for instance, is discontinuous:
// a fused multiply-add could flip a branch the reference does not take).
//
// Reference semantics restated (paths relative to the reference root):
//   gridworld/utils.py:9-43                                  to_scaled / to_raw
//   gridworld/agents/energy_storage/energy_storage_env.py:100-157
//   gridworld/agents/pv/pv_profile_env.py:102-148
//   gridworld/agents/vehicles/ev_charging_env.py:135-264
//   gridworld/agents/buildings/five_zone_rom_env.py:147-335
//   gridworld/agents/buildings/five_zone_rom_dynamics.py:12-114
//   gridworld/base.py:114-156                                composite agent
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/pgw.h"

#if defined(__CUDACC__)
#define PGW_HD __host__ __device__ __forceinline__
#else
#define PGW_HD inline
#endif
#define PGW_RESTRICT __restrict__

namespace pgw {

// Everything one (env, agent) worker needs; all per-env arrays are rows x E.
struct AgentIO {
  int E;
  // The arrays never alias each other; telling the compiler lets it overlap the many
  // independent load -> divide -> store chains of one agent (the kernel is latency bound).
  const double* PGW_RESTRICT actions;   // [act_dim][E]   (unused at reset)
  double* PGW_RESTRICT obs;             // [obs_dim][E]
  double* PGW_RESTRICT sd;              // [sd_rows][E]
  uint32_t* PGW_RESTRICT si;            // [si_rows][E]
  const double* PGW_RESTRICT init_soc;  // [num_storage][E] or nullptr (reset only)
  const double* PGW_RESTRICT vmin;      // [E]    lagged grid variables (previous solve) or nullptr
  const double* PGW_RESTRICT vmax;      // [E]
  const double* PGW_RESTRICT vbus;      // [A][E]
  const double* PGW_RESTRICT dpar;
  const int32_t* PGW_RESTRICT ipar;
  const double* PGW_RESTRICT drow;      // event row (doubles)
  const int32_t* PGW_RESTRICT irow;     // event row (int32)
};

PGW_HD double clip(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }

// x / d when r = RN(1/d) was computed on the host: one multiply plus one FMA-based
// correction (Markstein) returns the correctly rounded IEEE quotient, i.e. the same bits
// as the reference's division, at a third of the dependent latency of a full FP64
// divide.  fma() is explicit here; everything else in this file is compiled -fmad=false.
PGW_HD double div_by(double x, double d, double r) {
  const double q = x * r;
  const double rem = fma(-q, d, x);
  return fma(rem, r, q);
}

// utils.py:27-43
PGW_HD double to_raw(double y, double lo, double hi) {
  y = clip(y, -1.0, 1.0);
  return (y * (hi - lo) + (hi + lo)) / 2.0;
}
// utils.py:9-24; inv = RN(1 / (hi - lo))
PGW_HD double to_scaled(double x, double lo, double hi, double inv) {
  x = clip(x, lo, hi);
  return div_by(2.0 * x - (lo + hi), hi - lo, inv);
}

// ------------------------------------------------------------------ storage
PGW_HD double storage_obs(const pgw_component& c, const double* dp, double soc) {
  return (c.flags & PGW_F_RESCALE) ? to_scaled(soc, dp[0], dp[1], dp[7]) : soc;
}

PGW_HD void storage_reset(const pgw_component& c, const AgentIO& io, int e) {
  const double* dp = io.dpar + c.dpar_off;
  double* soc = io.sd + (size_t)c.sd_off * io.E + e;
  const int ord = io.ipar[c.ipar_off];
  const double init = io.init_soc != nullptr ? io.init_soc[(size_t)ord * io.E + e] : dp[6];
  *soc = clip(init, dp[0], dp[1]);                                     // :88-89
  io.obs[(size_t)c.obs_off * io.E + e] = storage_obs(c, dp, *soc);
}

PGW_HD void storage_step(const pgw_component& c, const AgentIO& io, int e, double& p_out) {
  const double* dp = io.dpar + c.dpar_off;
  const double lo = dp[0], hi = dp[1], eta_c = dp[2], eta_d = dp[3], pmax = dp[4], dt = dp[5];
  const double inv_eta_d = dp[8], inv_dt = dp[9];
  double a = io.actions[(size_t)c.act_off * io.E + e];
  if (c.flags & PGW_F_RESCALE) a = to_raw(a, -1.0, 1.0);
  double* soc_p = io.sd + (size_t)c.sd_off * io.E + e;
  double soc = *soc_p;
  double p = a * pmax;
  // validate_power :100-128 (the clamps omit the efficiencies, as in the reference)
  if (p > 0.0) {
    if (soc - div_by(p * dt, eta_d, inv_eta_d) < lo) p = div_by(fmax(soc - lo, 0.0), dt, inv_dt);
  } else if (p < 0.0) {
    if (soc - eta_c * p * dt > hi) p = div_by(-fmax(hi - soc, 0.0), dt, inv_dt);
  }
  if (p < 0.0) {
    soc -= eta_c * p * dt;
    soc = fmin(soc, hi);
  } else if (p > 0.0) {
    soc -= div_by(p * dt, eta_d, inv_eta_d);
    soc = fmax(soc, lo);
  }
  *soc_p = soc;
  io.obs[(size_t)c.obs_off * io.E + e] = storage_obs(c, dp, soc);
  p_out = -p;
}

// ------------------------------------------------------------------ PV
PGW_HD void pv_obs(const pgw_component& c, const AgentIO& io, int e, double raw_power) {
  const double* dp = io.dpar + c.dpar_off;
  const bool rs = (c.flags & PGW_F_RESCALE) != 0;
  io.obs[(size_t)c.obs_off * io.E + e] = rs ? to_scaled(raw_power, dp[0], dp[1], dp[4]) : raw_power;
  if (c.flags & PGW_F_GRID_AWARE) {
    const double v = io.vmin[e];
    io.obs[(size_t)(c.obs_off + 1) * io.E + e] = rs ? to_scaled(v, dp[2], dp[3], dp[5]) : v;
  }
}

PGW_HD void pv_step(const pgw_component& c, const AgentIO& io, int e, double& p_out,
                    double& rew) {
  double a = io.actions[(size_t)c.act_off * io.E + e];
  if (c.flags & PGW_F_RESCALE) a = to_raw(a, 0.0, 1.0);
  const double raw_power = -io.drow[c.dtab_off];      // obs of the PRE-increment row (:143)
  pv_obs(c, io, e, raw_power);
  p_out = a * raw_power;
  rew = 0.0;
  if (c.flags & PGW_F_PV_VOLT_REWARD) {               // heterogeneous.py:46-52 (lagged vmin)
    const double v = io.vmin[e];
    const double viol = fmin(0.0, v - 0.95) + fmin(0.0, 1.05 - v);
    const double s = 1000.0 * viol;
    rew = -(s * s);
  }
}

// ------------------------------------------------------------------ EV station
// One pass over the vehicles parked at the event's time (the roster is shared by
// all envs, so the window is a per-event list; only the "energy > 0" part of the
// reference's charging set is per-env).  kwh = energy one vehicle may take now.
PGW_HD void ev_advance(const pgw_component& c, const AgentIO& io, int e, double kwh,
                       double& p_out, double& rew) {
  const double* dp = io.dpar + c.dpar_off;
  const int32_t* ip = io.ipar + c.ipar_off;
  const int n = ip[0], words = ip[1], cap = ip[2];
  const double rate = dp[0], mult = dp[2];
  const double* obs_high = dp + 7;
  const double* inv_high = dp + 13;
  const double* end_park = dp + 21;
  const double t_now = io.drow[c.dtab_off], t_next = io.drow[c.dtab_off + 1];
  const int32_t* ir = io.irow + c.itab_off;
  const int n_win = ir[0], n_left = ir[1];
  const int32_t* win = ir + 2;
  const int32_t* left = ir + 2 + cap;
  double* energy = io.sd + (size_t)c.sd_off * io.E + e;
  uint32_t* mask = io.si + (size_t)c.si_off * io.E + e;

  double consumed = 0.0, demand = 0.0, deficit_sum = 0.0;
  int active = 0, n_deficit = 0;
  uint32_t word = 0;
  int cur_word = 0;
  for (int w = 0; w < words; ++w) mask[(size_t)w * io.E] = 0u;
  for (int k = 0; k < n_win; ++k) {                   // ascending vehicle index
    const int i = win[k];
    const double need = energy[(size_t)i * io.E];
    if (!(need > 0.0)) continue;                      // :191
    if ((i >> 5) != cur_word) {
      if (word) mask[(size_t)cur_word * io.E] = word;
      cur_word = i >> 5;
      word = 0;
    }
    word |= 1u << (i & 31);
    ++active;
    demand += need;                                   // :210
    const double left_h = div_by(end_park[i] - t_now, 60.0, dp[20]);
    if (left_h <= 0.0) continue;                      // :218-220
    deficit_sum += fmax(0.0, rate - need / left_h);   // :221-223
    ++n_deficit;
    const double delta = fmin(kwh, need);             // :226-228
    energy[(size_t)i * io.E] = need - delta;
    consumed += delta;
  }
  if (word) mask[(size_t)cur_word * io.E] = word;
  double unserved = 0.0;                              // :240-243 over window(k-1) \ window(k)
  for (int k = 0; k < n_left; ++k) unserved += energy[(size_t)left[k] * io.E];
  (void)n;

  const double s_consumed = mult * consumed;
  const double raw[6] = {t_next, mult * (double)active, s_consumed, mult * demand,
                         n_deficit == 0 ? 0.0 : deficit_sum / (double)n_deficit, unserved};
  const bool rs = (c.flags & PGW_F_RESCALE) != 0;
#pragma unroll
  for (int j = 0; j < 6; ++j)
    io.obs[(size_t)(c.obs_off + j) * io.E + e] =
        rs ? to_scaled(raw[j], 0.0, obs_high[j], inv_high[j]) : raw[j];
  p_out = s_consumed;                                 // kWh per step reported as kW (:255)
  const double over = fmax(0.0, s_consumed - dp[5]);  // :135-142
  rew = div_by(-dp[3] * (unserved * unserved) + -dp[4] * (over * over), dp[6], dp[19]);
}

PGW_HD void ev_step(const pgw_component& c, const AgentIO& io, int e, double& p_out,
                    double& rew) {
  const double* dp = io.dpar + c.dpar_off;
  double a = io.actions[(size_t)c.act_off * io.E + e];
  if (c.flags & PGW_F_RESCALE) a = to_raw(a, 0.0, 1.0);
  ev_advance(c, io, e, (a * dp[0]) * dp[1], p_out, rew);      // :182-183
}

PGW_HD void ev_reset(const pgw_component& c, const AgentIO& io, int e) {
  const double* dp = io.dpar + c.dpar_off;
  const int n = io.ipar[c.ipar_off];
  const double* e0 = dp + 21 + n;
  double* energy = io.sd + (size_t)c.sd_off * io.E + e;
  for (int i = 0; i < n; ++i) energy[(size_t)i * io.E] = e0[i];
  // Hidden step with action=None -> _action_space.low = 0 (:163, :178).  With
  // rescale_spaces the reference still pushes that 0 through to_raw, i.e. 0.5.
  double a = 0.0, p, r;
  if (c.flags & PGW_F_RESCALE) a = to_raw(a, 0.0, 1.0);
  ev_advance(c, io, e, (a * dp[0]) * dp[1], p, r);
}

// ------------------------------------------------------------------ five-zone building
struct BuildingPar {
  const double *A, *B, *C, *K, *mean, *Tinit, *low, *high, *inv;
  double w_energy, w_comfort;
  const int32_t *sel, *nbr;
  uint32_t obs_mask;
};

PGW_HD BuildingPar building_par(const pgw_component& c, const AgentIO& io) {
  const double* dp = io.dpar + c.dpar_off;
  const int32_t* ip = io.ipar + c.ipar_off;
  BuildingPar b;
  b.A = dp; b.B = dp + 5; b.C = dp + 25; b.K = dp + 30; b.mean = dp + 35; b.Tinit = dp + 40;
  b.w_energy = dp[45]; b.w_comfort = dp[46];
  b.low = dp + 47; b.high = dp + 47 + c.obs_dim; b.inv = dp + 47 + 2 * c.obs_dim;
  b.sel = ip; b.nbr = ip + 20; b.obs_mask = (uint32_t)ip[40];
  return b;
}

// build_u_vector (dynamics.py:12-41): candidate inputs, then the model's selection of 4
PGW_HD void building_inputs(const BuildingPar& b, const double T[5], double t_oa,
                            const double* q_solar, const double* q_x, bool use_q_cool,
                            const double* flow, double t_dis, double u[5][4]) {
#pragma unroll
  for (int z = 0; z < 5; ++z) {
    double cand[8];
    cand[0] = t_oa - T[z];
    cand[1] = q_solar[z];
    cand[2] = use_q_cool ? 0.0 : q_x[z];              // q_int (never selected by the shipped model)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int y = b.nbr[z * 4 + i];
      double ty = T[0];
#pragma unroll
      for (int q = 1; q < 5; ++q) ty = (y == q) ? T[q] : ty;
      cand[3 + i] = ty - T[z];
    }
    cand[7] = use_q_cool ? q_x[z] : flow[z] * (t_dis - T[z]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int s = b.sel[z * 4 + j];
      double v = cand[0];
#pragma unroll
      for (int q = 1; q < 8; ++q) v = (s == q) ? cand[q] : v;
      u[z][j] = v;
    }
  }
}

// state_update (dynamics.py:44-55); B already rounded through float32 on the host
PGW_HD double building_x_next(const BuildingPar& b, int z, double x, const double u[4]) {
  const double* B = b.B + z * 4;
  return b.A[z] * x + (((B[0] * u[0] + B[1] * u[1]) + B[2] * u[2]) + B[3] * u[3]);
}

// FiveZoneROMThermalEnergyEnv.step_reward (five_zone_rom_env.py:315-335)
PGW_HD double building_reward(const BuildingPar& b, const double T[5], double lb, double ub,
                              double p_consumed) {
  const double energy = div_by(-p_consumed, 12.0, 1.0 / 12.0);
  double comfort = 0.0;
#pragma unroll
  for (int z = 0; z < 5; ++z) {
    const double err = fmax(fmax(T[z] - ub, lb - T[z]), 0.0);
    comfort += err * err;
  }
  comfort = -comfort;
  return b.w_energy * energy + b.w_comfort * comfort;
}

// get_obs (five_zone_rom_env.py:228-283): values in state-dict order, bounds in label order
PGW_HD void building_obs(const pgw_component& c, const BuildingPar& b, const AgentIO& io,
                         int e, const double T[5], double lb, double ub, double t_oa,
                         double p_consumed, double tod) {
  const bool rs = (c.flags & PGW_F_RESCALE) != 0;
  const int a = c.agent;
  int slot = 0;
  auto put = [&](int src, double v) {
    if (b.obs_mask & (1u << src)) {
      double o = clip(v, b.low[slot], b.high[slot]);
      if (rs) o = to_scaled(o, b.low[slot], b.high[slot], b.inv[slot]);
      io.obs[(size_t)(c.obs_off + slot) * io.E + e] = o;
      ++slot;
    }
  };
#pragma unroll
  for (int z = 0; z < 5; ++z) put(z, T[z]);
#pragma unroll
  for (int z = 0; z < 5; ++z) put(5 + z, T[z] - ub);
#pragma unroll
  for (int z = 0; z < 5; ++z) put(10 + z, lb - T[z]);
  put(15, lb);
  put(16, ub);
  put(17, t_oa);
  put(18, p_consumed);
  put(19, tod);
  if (b.obs_mask & (7u << 20)) {
    put(20, io.vbus[(size_t)a * io.E + e]);
    put(21, io.vmin[e]);
    put(22, io.vmax[e]);
  }
  put(23, INFINITY);                                   // p_setpoint default (:268)
}

PGW_HD void building_reset(const pgw_component& c, const AgentIO& io, int e) {
  const BuildingPar b = building_par(c, io);
  const double* row = io.drow + c.dtab_off;
  double* sd = io.sd + (size_t)c.sd_off * io.E + e;
  double x[5], T[5], u[5][4];
#pragma unroll
  for (int z = 0; z < 5; ++z) { x[z] = sd[(size_t)z * io.E]; T[z] = b.Tinit[z]; }   // x persists (:94)
  building_inputs(b, T, row[0], row + 1, row + 6, true, nullptr, 0.0, u);
  for (int rep = 0; rep < 2; ++rep) {                  // filter_update x2 (dynamics.py:58-72)
#pragma unroll
    for (int z = 0; z < 5; ++z) {
      x[z] = building_x_next(b, z, x[z], u[z]);
      x[z] += b.K[z] * ((T[z] - b.mean[z]) - b.C[z] * x[z]);
    }
  }
#pragma unroll
  for (int z = 0; z < 5; ++z) {
    T[z] = b.C[z] * x[z] + b.mean[z];
    sd[(size_t)z * io.E] = x[z];
  }
  sd[(size_t)5 * io.E] = 0.0;                          // p_consumed
  building_obs(c, b, io, e, T, row[12], row[13], row[11], 0.0, row[14]);
}

PGW_HD void building_step(const pgw_component& c, const AgentIO& io, int e, double& p_out,
                          double& rew) {
  const double kLow[6] = {0.22, 0.22, 0.22, 0.22, 0.32, 10.0};   // :22-26
  const double kHigh[6] = {2.2, 2.2, 2.2, 2.2, 3.2, 16.0};
  const BuildingPar b = building_par(c, io);
  const double* row = io.drow + c.dtab_off;
  double* sd = io.sd + (size_t)c.sd_off * io.E + e;
  double act[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const double a = io.actions[(size_t)(c.act_off + i) * io.E + e];
    act[i] = (c.flags & PGW_F_RESCALE) ? to_raw(a, kLow[i], kHigh[i]) : a;
  }
  double x[5], T[5], u[5][4];
#pragma unroll
  for (int z = 0; z < 5; ++z) { x[z] = sd[(size_t)z * io.E]; T[z] = b.C[z] * x[z] + b.mean[z]; }
  if (c.flags & PGW_F_STALE_REWARD)                    // stand-alone agent: pre-step state (:215)
    rew = building_reward(b, T, row[15], row[16], sd[(size_t)5 * io.E]);
  const double t_oa = row[0];
  building_inputs(b, T, t_oa, row + 1, row + 6, false, act, act[5], u);
#pragma unroll
  for (int z = 0; z < 5; ++z) {
    x[z] = building_x_next(b, z, x[z], u[z]);
    T[z] = b.C[z] * x[z] + b.mean[z];
    sd[(size_t)z * io.E] = x[z];
  }
  const double flow = (((act[0] + act[1]) + act[2]) + act[3]) + act[4];
  // flow**3: the reference calls libm pow; x*x*x differs from it by at most 1 ulp
  const double p = (0.0076 * ((flow * flow) * flow) + 4.8865) + fmax(0.0, flow * (t_oa - act[5]));
  sd[(size_t)5 * io.E] = p;
  building_obs(c, b, io, e, T, row[12], row[13], row[11], p, row[14]);
  if (!(c.flags & PGW_F_STALE_REWARD)) rew = building_reward(b, T, row[12], row[13], p);
  p_out = p;
}

// ------------------------------------------------------------------ one agent of one env
// MultiComponentEnv.step (base.py:114-139): components in order, real power summed,
// reward = sum of the components' post-step rewards.  A single-component agent is the
// one-element case (its own step reward, multiagent_env.py:168).
PGW_HD void agent_step(const pgw_agent& ag, const pgw_component* comps, const AgentIO& io, int e,
                       double& p_agent, double& r_agent) {
  p_agent = 0.0;
  r_agent = 0.0;
  for (int ci = ag.comp_begin; ci < ag.comp_end; ++ci) {
    const pgw_component c = comps[ci];
    double p = 0.0, r = 0.0;
    switch (c.type) {
      case PGW_STORAGE: storage_step(c, io, e, p); break;
      case PGW_PV: pv_step(c, io, e, p, r); break;
      case PGW_EV: ev_step(c, io, e, p, r); break;
      case PGW_BUILDING: building_step(c, io, e, p, r); break;
      default: break;
    }
    p_agent += p;
    r_agent += r;
  }
}

PGW_HD void agent_reset(const pgw_agent& ag, const pgw_component* comps, const AgentIO& io, int e) {
  for (int ci = ag.comp_begin; ci < ag.comp_end; ++ci) {
    const pgw_component c = comps[ci];
    switch (c.type) {
      case PGW_STORAGE: storage_reset(c, io, e); break;
      case PGW_PV: pv_obs(c, io, e, -io.drow[c.dtab_off]); break;
      case PGW_EV: ev_reset(c, io, e); break;
      case PGW_BUILDING: building_reset(c, io, e); break;
      default: break;
    }
  }
}

}  // namespace pgw
