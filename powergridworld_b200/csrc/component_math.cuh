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
#if defined(__CUDACC__)
#define PGW_NO_UNROLL _Pragma("unroll 1")
#else
#define PGW_NO_UNROLL
#endif

namespace pgw {

// Per-thread scratch array: element i lives at p[i * stride].
struct Scratch {
  double* p;
  int stride;
  PGW_HD double& operator[](int i) const { return p[(size_t)i * stride]; }
};
constexpr int kScratchDoubles = 35;

// Everything one (env, agent) worker needs; all per-env arrays are rows x E.
struct AgentIO {
  Scratch scr;
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
  // Ascending vehicle index, four vehicles per trip: their energies are loaded up front so
  // that four independent HBM/L2 requests are in flight per thread (a vehicle's store can
  // never alias another vehicle's load, which the compiler cannot know).
  for (int k0 = 0; k0 < n_win; k0 += 4) {
    int idx[4];
    double need4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      idx[j] = k0 + j < n_win ? win[k0 + j] : -1;
      need4[j] = idx[j] >= 0 ? energy[(size_t)idx[j] * io.E] : 0.0;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = idx[j];
      const double need = need4[j];
      if (i < 0 || !(need > 0.0)) continue;             // :191
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
// Rolled loops over zones and observation slots: the per-thread arrays they index live in
// a scratch area (shared memory on the GPU, strided by the CTA size so lanes hit distinct
// banks; a plain local array in the host build).  Keeps the kernel's code small -- the
// fully unrolled form was 140 kB of SASS and stalled on instruction-cache misses.
enum { BSCR_T = 0, BSCR_ACT = 5, BSCR_SRC = 11, BSCR_SIZE = 35 };   // T[5], act[6], src[24]
enum { U_OUTDOOR = 0, U_SOLAR = 1, U_INTERNAL = 2, U_NEIGHBOR = 3, U_COOLING = 4 };

struct BuildingPar {
  const double *A, *B, *C, *K, *mean, *Tinit, *low, *high, *inv;
  double w_energy, w_comfort;
  const int32_t *u_kind, *u_arg, *obs_src;
};

PGW_HD BuildingPar building_par(const pgw_component& c, const AgentIO& io) {
  const double* dp = io.dpar + c.dpar_off;
  const int32_t* ip = io.ipar + c.ipar_off;
  BuildingPar b;
  b.A = dp; b.B = dp + 5; b.C = dp + 25; b.K = dp + 30; b.mean = dp + 35; b.Tinit = dp + 40;
  b.w_energy = dp[45]; b.w_comfort = dp[46];
  b.low = dp + 47; b.high = dp + 47 + c.obs_dim; b.inv = dp + 47 + 2 * c.obs_dim;
  b.u_kind = ip; b.u_arg = ip + 20; b.obs_src = ip + 40;
  return b;
}

// One selected model input of zone z (build_u_vector, dynamics.py:12-41; the host resolved
// the model's input_sel_list / neighbors into a kind + argument per input).
PGW_HD double building_input(const BuildingPar& b, const Scratch& scr, int z, int j, double t_oa,
                             const double* q_solar, const double* q_x, bool at_reset) {
  const int kind = b.u_kind[z * 4 + j];
  const double Tz = scr[BSCR_T + z];
  switch (kind) {
    case U_OUTDOOR: return t_oa - Tz;
    case U_SOLAR: return q_solar[z];
    case U_INTERNAL: return at_reset ? 0.0 : q_x[z];
    case U_NEIGHBOR: return scr[BSCR_T + b.u_arg[z * 4 + j]] - Tz;
    default:  // U_COOLING: q_cool at reset, m_dot (T_discharge - T_z) afterwards
      return at_reset ? q_x[z] : scr[BSCR_ACT + z] * (scr[BSCR_ACT + 5] - Tz);
  }
}

// state_update (dynamics.py:44-55); B already rounded through float32 on the host
PGW_HD double building_x_next(const BuildingPar& b, int z, double x, double u0, double u1,
                              double u2, double u3) {
  const double* B = b.B + z * 4;
  return b.A[z] * x + (((B[0] * u0 + B[1] * u1) + B[2] * u2) + B[3] * u3);
}

// FiveZoneROMThermalEnergyEnv.step_reward (five_zone_rom_env.py:315-335) on temps T[0..5)
PGW_HD double building_reward(const BuildingPar& b, const Scratch& scr, int t_off, double lb,
                              double ub, double p_consumed) {
  const double energy = div_by(-p_consumed, 12.0, 1.0 / 12.0);
  double comfort = 0.0;
  PGW_NO_UNROLL
  for (int z = 0; z < 5; ++z) {
    const double T = scr[t_off + z];
    const double err = fmax(fmax(T - ub, lb - T), 0.0);
    comfort += err * err;
  }
  comfort = -comfort;
  return b.w_energy * energy + b.w_comfort * comfort;
}

// get_obs (five_zone_rom_env.py:228-283): the 24 possible sources in state-dict order, then
// the selected ones (obs_src, in that same order) against the bounds in label order.
PGW_HD void building_obs(const pgw_component& c, const BuildingPar& b, const AgentIO& io,
                         int e, const Scratch& scr, double lb, double ub, double t_oa,
                         double p_consumed, double tod, bool grid) {
  PGW_NO_UNROLL
  for (int z = 0; z < 5; ++z) {
    const double T = scr[BSCR_T + z];
    scr[BSCR_SRC + z] = T;
    scr[BSCR_SRC + 5 + z] = T - ub;
    scr[BSCR_SRC + 10 + z] = lb - T;
  }
  scr[BSCR_SRC + 15] = lb;
  scr[BSCR_SRC + 16] = ub;
  scr[BSCR_SRC + 17] = t_oa;
  scr[BSCR_SRC + 18] = p_consumed;
  scr[BSCR_SRC + 19] = tod;
  if (grid) {
    scr[BSCR_SRC + 20] = io.vbus[(size_t)c.agent * io.E + e];
    scr[BSCR_SRC + 21] = io.vmin[e];
    scr[BSCR_SRC + 22] = io.vmax[e];
  }
  scr[BSCR_SRC + 23] = INFINITY;                        // p_setpoint default (:268)
  const bool rs = (c.flags & PGW_F_RESCALE) != 0;
  PGW_NO_UNROLL
  for (int slot = 0; slot < c.obs_dim; ++slot) {
    const double lo = b.low[slot], hi = b.high[slot];
    double o = clip(scr[BSCR_SRC + b.obs_src[slot]], lo, hi);   // np.clip; to_scaled's own clip
    if (rs) o = div_by(2.0 * o - (lo + hi), hi - lo, b.inv[slot]);   // is then the identity
    io.obs[(size_t)(c.obs_off + slot) * io.E + e] = o;
  }
}

PGW_HD void building_reset(const pgw_component& c, const AgentIO& io, int e) {
  const BuildingPar b = building_par(c, io);
  const Scratch& scr = io.scr;
  const double* row = io.drow + c.dtab_off;
  double* sd = io.sd + (size_t)c.sd_off * io.E + e;
  for (int z = 0; z < 5; ++z) scr[BSCR_T + z] = b.Tinit[z];
  PGW_NO_UNROLL
  for (int z = 0; z < 5; ++z) {
    // the inputs are built once from T_init; two filter updates (dynamics.py:58-72)
    const double u0 = building_input(b, scr, z, 0, row[0], row + 1, row + 6, true);
    const double u1 = building_input(b, scr, z, 1, row[0], row + 1, row + 6, true);
    const double u2 = building_input(b, scr, z, 2, row[0], row + 1, row + 6, true);
    const double u3 = building_input(b, scr, z, 3, row[0], row + 1, row + 6, true);
    double x = sd[(size_t)z * io.E];                    // x persists across resets (:94)
    for (int rep = 0; rep < 2; ++rep) {
      x = building_x_next(b, z, x, u0, u1, u2, u3);
      x += b.K[z] * ((b.Tinit[z] - b.mean[z]) - b.C[z] * x);
    }
    sd[(size_t)z * io.E] = x;
    scr[BSCR_ACT + z] = b.C[z] * x + b.mean[z];         // new temps, parked until all zones are done
  }
  for (int z = 0; z < 5; ++z) scr[BSCR_T + z] = scr[BSCR_ACT + z];
  sd[(size_t)5 * io.E] = 0.0;                           // p_consumed
  building_obs(c, b, io, e, scr, row[12], row[13], row[11], 0.0, row[14],
               (c.flags & PGW_F_GRID_AWARE) != 0);
}

PGW_HD void building_step(const pgw_component& c, const AgentIO& io, int e, double& p_out,
                          double& rew) {
  const BuildingPar b = building_par(c, io);
  const Scratch& scr = io.scr;
  const double* row = io.drow + c.dtab_off;
  double* sd = io.sd + (size_t)c.sd_off * io.E + e;
  const bool rs = (c.flags & PGW_F_RESCALE) != 0;
  double flow = 0.0;
  PGW_NO_UNROLL
  for (int i = 0; i < 6; ++i) {                         // action bounds :22-26
    const double lo = i < 4 ? 0.22 : (i == 4 ? 0.32 : 10.0);
    const double hi = i < 4 ? 2.2 : (i == 4 ? 3.2 : 16.0);
    const double a = io.actions[(size_t)(c.act_off + i) * io.E + e];
    const double raw = rs ? to_raw(a, lo, hi) : a;
    scr[BSCR_ACT + i] = raw;
    if (i < 5) flow = i == 0 ? raw : flow + raw;
  }
  for (int z = 0; z < 5; ++z) scr[BSCR_T + z] = b.C[z] * sd[(size_t)z * io.E] + b.mean[z];
  if (c.flags & PGW_F_STALE_REWARD)                     // stand-alone agent: pre-step state (:215)
    rew = building_reward(b, scr, BSCR_T, row[15], row[16], sd[(size_t)5 * io.E]);
  const double t_oa = row[0], t_dis = scr[BSCR_ACT + 5];
  PGW_NO_UNROLL
  for (int z = 0; z < 5; ++z) {                         // all inputs use the OLD temperatures
    const double u0 = building_input(b, scr, z, 0, t_oa, row + 1, row + 6, false);
    const double u1 = building_input(b, scr, z, 1, t_oa, row + 1, row + 6, false);
    const double u2 = building_input(b, scr, z, 2, t_oa, row + 1, row + 6, false);
    const double u3 = building_input(b, scr, z, 3, t_oa, row + 1, row + 6, false);
    const double x = building_x_next(b, z, sd[(size_t)z * io.E], u0, u1, u2, u3);
    sd[(size_t)z * io.E] = x;
    scr[BSCR_SRC + z] = b.C[z] * x + b.mean[z];         // new temps, parked in the source area
  }
  for (int z = 0; z < 5; ++z) scr[BSCR_T + z] = scr[BSCR_SRC + z];
  // flow**3: the reference calls libm pow; x*x*x differs from it by at most 1 ulp
  const double p = (0.0076 * ((flow * flow) * flow) + 4.8865) + fmax(0.0, flow * (t_oa - t_dis));
  sd[(size_t)5 * io.E] = p;
  building_obs(c, b, io, e, scr, row[12], row[13], row[11], p, row[14],
               (c.flags & PGW_F_GRID_AWARE) != 0);
  if (!(c.flags & PGW_F_STALE_REWARD)) rew = building_reward(b, scr, BSCR_T, row[12], row[13], p);
  p_out = p;
}

// ---- fast path: the reference's shipped model + default observation set ----------------
// PGW_F_BUILDING_FAST is set by the host when (a) every zone's model inputs are
// [outdoor, cooling, neighbour, solar] (the input_sel_list of state_space_model.p) and
// (b) the observation set is defaults.obs_config (upper/lower violation per zone, comfort
// band, outdoor temperature, p_consumed, time of day).  Same arithmetic, same order of
// operations as the generic path, but straight-line code on registers: the table-driven
// form spends ~85 % of its instructions on indexing rather than on float64 math.
PGW_HD double pick5(const double T[5], int i) {
  double v = T[0];
  v = i == 1 ? T[1] : v;
  v = i == 2 ? T[2] : v;
  v = i == 3 ? T[3] : v;
  v = i == 4 ? T[4] : v;
  return v;
}

PGW_HD void building_step_fast(const pgw_component& c, const AgentIO& io, int e, double& p_out,
                               double& rew) {
  const double* dp = io.dpar + c.dpar_off;
  const int32_t* u_arg = io.ipar + c.ipar_off + 20;
  const double *A = dp, *B = dp + 5, *C = dp + 25, *mean = dp + 35;
  const double *low = dp + 47, *high = dp + 62, *inv = dp + 77;     // obs_dim == 15
  const double* row = io.drow + c.dtab_off;
  const size_t E = (size_t)io.E;
  const bool rs = (c.flags & PGW_F_RESCALE) != 0;
  const double kLow[6] = {0.22, 0.22, 0.22, 0.22, 0.32, 10.0};      // action bounds :22-26
  const double kHigh[6] = {2.2, 2.2, 2.2, 2.2, 3.2, 16.0};

  const double* ap = io.actions + (size_t)c.act_off * E + e;
  double* sp = io.sd + (size_t)c.sd_off * E + e;
  double act[6], x[5], T[5], Tn[5];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const double a = ap[i * E];
    act[i] = rs ? to_raw(a, kLow[i], kHigh[i]) : a;
  }
#pragma unroll
  for (int z = 0; z < 5; ++z) {
    x[z] = sp[z * E];
    T[z] = C[z] * x[z] + mean[z];
  }
  const double t_oa = row[0], t_dis = act[5];
#pragma unroll
  for (int z = 0; z < 5; ++z) {                  // inputs [outdoor, cooling, neighbour, solar]
    const double u0 = t_oa - T[z];
    const double u1 = act[z] * (t_dis - T[z]);
    const double u2 = pick5(T, u_arg[z * 4 + 2]) - T[z];
    const double u3 = row[1 + z];
    const double* Bz = B + z * 4;
    const double xn = A[z] * x[z] + (((Bz[0] * u0 + Bz[1] * u1) + Bz[2] * u2) + Bz[3] * u3);
    sp[z * E] = xn;
    Tn[z] = C[z] * xn + mean[z];
  }
  const double flow = (((act[0] + act[1]) + act[2]) + act[3]) + act[4];
  const double p = (0.0076 * ((flow * flow) * flow) + 4.8865) + fmax(0.0, flow * (t_oa - t_dis));
  sp[5 * E] = p;

  const double lb = row[12], ub = row[13];
  double* op = io.obs + (size_t)c.obs_off * E + e;
  double comfort = 0.0;
#pragma unroll
  for (int z = 0; z < 5; ++z) {
    const double up = Tn[z] - ub, lo = lb - Tn[z];
    double o = clip(up, low[z], high[z]);
    op[z * E] = rs ? div_by(2.0 * o - (low[z] + high[z]), high[z] - low[z], inv[z]) : o;
    o = clip(lo, low[5 + z], high[5 + z]);
    op[(5 + z) * E] =
        rs ? div_by(2.0 * o - (low[5 + z] + high[5 + z]), high[5 + z] - low[5 + z], inv[5 + z]) : o;
    const double err = fmax(fmax(up, lo), 0.0);
    comfort += err * err;
  }
  const double tail[5] = {lb, ub, row[11], p, row[14]};
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const int s = 10 + j;
    const double o = clip(tail[j], low[s], high[s]);
    op[s * E] = rs ? div_by(2.0 * o - (low[s] + high[s]), high[s] - low[s], inv[s]) : o;
  }
  const double energy = div_by(-p, 12.0, 1.0 / 12.0);
  rew = dp[45] * energy + dp[46] * (-comfort);
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
      case PGW_BUILDING:
        if (c.flags & PGW_F_BUILDING_FAST) building_step_fast(c, io, e, p, r);
        else building_step(c, io, e, p, r);
        break;
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
