// Internal launch-parameter blocks shared between api.cu and the kernel files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pgw.h"

namespace pgw {

struct CompParams {
  int E, A;
  int event_mode;          // 0 = reset (event row 0), 1 = step (event row clock+1)
  int advance_clock;       // 1 when this kernel is the last one of the step (no feeder)
  const pgw_agent* agents;
  const pgw_component* comps;
  const double* dpar;
  const int32_t* ipar;
  const double* dtab;
  const int32_t* itab;
  int dstride, istride;
  const double* actions;
  double* obs;
  double* rew;
  uint8_t* done;
  double* sd;
  uint32_t* si;
  const double* init_soc;
  const double* vmin;
  const double* vmax;
  const double* vbus;
  double* agent_p;
  double* ep_ret;
  int* clock;
  unsigned int* ticket;
};

struct PfParams {
  int E, A, nb, nn, nl, nbp, nnp, max_iter;
  double tol;
  int event_mode;          // 0 = reset (base load only, event row 0), 1 = step
  int advance_clock;
  const double2* zbbT;     // [nb][nbp]   zbbT[j*nbp+k] = Zbb[k][j]
  const double2* u0;       // [nbp]
  const double2* znbT;     // [nb][nnp]   znbT[k*nnp+n] = Znb[n][k]
  const double2* w;        // [nnp]
  const int32_t* branch_load;   // [nbp]
  const double* branch_share;   // [nbp]
  const int32_t* branch_model;  // [nbp]
  const double* vminpu;         // [nbp]
  const double* vmaxpu;         // [nbp]
  const pgw_agent* agents;
  const double* agent_p;   // [A][E]
  const double* load_kw;   // [nl][E] stand-alone solve: total kW per load (else nullptr)
  const double* load_kvar; // [nl][E]
  const double* dtab;
  int dstride;
  double* vmag;            // [nn][E]
  double* vmin;
  double* vmax;
  double* vbus;            // [A][E]
  int32_t* iters;          // [E]
  double* rew;             // [A][E] (step only)
  double* ep_ret;          // [A][E]
  double* viol;            // [E] voltage violation at the penalty node
  int penalty_node;
  double pvlo, pvhi, punit;
  int* clock;
  unsigned int* ticket;
};

struct StatsParams {
  int E, A, nn;
  const double* rew;
  const double* ep_ret;
  const double* viol;
  const int32_t* iters;
  const double* vmin;
  const double* vmax;
  const int* clock;
  double* out;             // [PGW_NUM_STATS]
};

cudaError_t launch_components(const CompParams& p, int smem_bytes, cudaStream_t s);
cudaError_t launch_powerflow(const PfParams& p, cudaStream_t s);
cudaError_t launch_stats(const StatsParams& p, cudaStream_t s);

}  // namespace pgw
