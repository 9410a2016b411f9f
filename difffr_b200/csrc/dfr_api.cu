// C ABI (include/dfr.h) of the B200-native DiffDFSPH step: context management, HBM layout,
// launch sequences.  Host code here only enqueues kernels; the step itself (solver loops, CFL,
// rigid update, chain rule) runs on the device (see dfr_kernels.cuh, dfr_rigid.cuh).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dfr.h"
#include "dfr_kernels.cuh"
#include "dfr_rigid.cuh"
#include "dfr_contact.cuh"
#include "dfr_slab.cuh"

#include <dlfcn.h>
#include <nccl.h>

using namespace dfr;

namespace {

struct HostBody {
  std::vector<double> x_local;  // n*3
  int64_t n = 0;
  int dynamic = 0;
  double density = 1000.0;
  double pos[3], q[4];
  double init_v[3] = {0, 0, 0}, init_w[3] = {0, 0, 0};
  int p_begin = 0;  // slice in the concatenated (unsorted) boundary layout
  int blk_begin = 0, blk_count = 0;  // accumulator rows of the boundary-side kernel
  BodyDev dev0;     // initial device record (reset() uploads it)
  bool init_staged = false;  // dfr_set_init_v_omega values waiting in the pinned staging buffer
};

template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  cudaError_t alloc(size_t count) {
    free();
    n = count;
    if (count == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
    // The clear runs on the legacy default stream and may return before the device has done it; the context's stream
    // is non-blocking, i.e. not ordered against that stream, so wait here - a kernel enqueued right after a (re)allocation
    // must not race with the clear (allocations only happen at finalize and when a capacity grows)
    if (e == cudaSuccess) e = cudaMemset(p, 0, count * sizeof(T));
    if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy);
    return e;
  }
  void free() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

}  // namespace

struct dfr_context {
  dfr_config cfg;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t ls = nullptr;  // the stream kernels are launched on: `stream`, or a capture stream while a step graph is recorded
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // CUDA-graph stepping: one instantiated graph per (gated?, buffer parity); see capture_step_graph
  struct StepGraph {
    cudaGraphExec_t exec = nullptr;
    cudaGraph_t graph = nullptr;
    int64_t n_static = 0, n_div_body = 0, n_prs_body = 0;  // kernels per step / per solver iteration (launch accounting)
  } step_graph[2][2];
  bool capturing = false;
  int graph_broken = 0;                 // a capture failed: stay on the stream path for good
  cudaStream_t cap_stream[3] = {nullptr, nullptr, nullptr};
  int cap_depth = 0;
  unsigned long long cap_gate_handle = 0;
  int64_t *cap_counter = nullptr;       // which of the StepGraph counters LAUNCH adds to while capturing
  StepGraph *cap_target = nullptr;
  std::string err;
  bool finalized = false;

  Params P;
  DevBuf<StepState> dSt;
  StepState *hSt = nullptr;  // pinned mirror

  // host-side scene description
  std::vector<double> h_fx, h_fv;
  int64_t nf0 = 0, nf_cap = 0;  // nf0: fluid particles of the scene
  int64_t nf_loc0 = 0;          // of those, held by this context at t = 0 (all of them unless slab-decomposed)
  std::vector<int> h_ids0;      // their ids (slab mode)
  std::vector<HostBody> bodies;
  std::vector<EmitterDev> h_emitters;
  DevBuf<EmitterDev> dEmitters;
  int n_static_p = 0, n_dyn_p = 0, n_b = 0, dyn_begin = 0;
  bool has_emitters = false;

  // fluid, cell order, double buffered where the per-step re-sort needs it
  DevBuf<double4> pos[2], vel[2];
  DevBuf<double> kappa[2], kappav[2];
  DevBuf<int> pid[2], pstate[2];
  int cur = 0, vcur = 0;  // current buffer of (pos,kappa,kappav,pid,pstate) / of vel
  DevBuf<double4> acc, sgp, normal, xk, xrho;  // xk = (x, stiffness), xrho = (x, density): single-record gathers
  DevBuf<double> density, factor, dadv, partials;
  // initial state in HBM (id order) restored by dfr_reset
  DevBuf<double4> pos_init, vel_init;
  DevBuf<double> kappa_init, kappav_init;

  // boundary particles: [static (cell order) | dynamic (body order)]
  DevBuf<double4> bpos, bvel, bx0, bpos_tmp, bx0_tmp;
  DevBuf<int> bbody, borig, bbody_tmp, borig_tmp;
  DevBuf<double> bvol;
  std::vector<int> h_borig;  // after the static sort

  DevBuf<BodyDev> dBodies;
  BodyDev *h_bodies = nullptr;  // pinned mirror of dBodies, refreshed with the state read-back that ends every dfr_step
  bool bodies_mirrored = false;
  double *h_init = nullptr;     // pinned staging of (init_v, init_omega) per body: uploaded at the start of the next step
  bool init_dirty = false;
  bool slab_needs_p2p_setup = false;
  DevBuf<MgrBlock> dMgr;
  DevBuf<double> acc_rows;
  DevBuf<int> blk_body, blk_first;
  int n_acc_blocks = 0;

  // grids
  DevBuf<unsigned int> cell_start_f, cell_start_s, cell_start_d, tile_sums, tile_sums_side;
  // recorded steps: the grid of the dynamic boundary particles is built on a side branch of the graph, next to the sort of the fluid
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool on_side_branch = false;
  DevBuf<unsigned char> near_s;  // per cell: a static boundary particle may be within reach (k_mark_near, marked once)
  DevBuf<unsigned char> near_d;  // the same for the dynamic boundary particles, re-marked with their cell table
  DevBuf<int> cell_of_p, rank_in_cell, sorted_src_f, sorted_src_d, cell_of_b, rank_b;
  // neighbour lists
  DevBuf<int> cnt_f, cnt_b, idx_f, idx_b, idx_d;
  DevBuf<unsigned int> off_d;
  int cap_f = 0, cap_b = 0;  // ELL row capacities (neighbours per particle)
  unsigned int cap_d = 0;

  // penalty rigid-rigid contact (dfr_contact.cuh)
  ContactParams CP;
  DevBuf<double> c_vol0, c_dens0, c_dens, c_records;
  DevBuf<double4> c_vel;
  DevBuf<int> c_order;
  std::vector<int> h_order;         // per dynamic body: particles (relative to dyn_begin) in the reference's storage order
  std::vector<double4> h_dynpos;    // scratch for the z-sort keys
  long long sort_counter = 0;       // TimeStepDiffDFSPH::m_counter
  bool contact_ready = false;

  // slab decomposition over several GPUs (dfr_slab.cuh); NCCL is loaded at run time so that single-GPU use needs none
  struct Slab {
    bool on = false;
    int rank = 0, n = 1;
    ncclComm_t comm = nullptr;
    SlabGeom G;
    std::vector<int> own0;               // ids owned at t = 0
    int h_ranges[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int send_cap = 0;
    DevBuf<double4> s_pos[2], s_vel[2], s_misc[2], r_misc;
    DevBuf<int> counts;                  // [0..1] my export counts, [2..3] what the neighbours export to me
    int *h_counts = nullptr;             // pinned, 4 ints
    double *h_stage = nullptr;           // pinned staging of dfr_load_fluid_state (10 doubles per local particle)
    size_t h_stage_n = 0;
    DevBuf<double> body_buf;
    long long exchanged_bytes = 0;       // NVLink traffic of the steps since reset (both directions, this rank)
    // peer-memory transport of the ghost updates (cudaIpc-mapped neighbour buffers, NVLink stores from the producing
    // kernels + one flag per pass); falls back to NCCL send/recv when the mapping is not available
    bool p2p = false;
    double4 *peer_lo[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}, *peer_hi[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    unsigned long long *peer_lo_flags = nullptr, *peer_hi_flags = nullptr;
    void *ipc_opened[14] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    DevBuf<unsigned long long> flags;    // [0] raised by the low neighbour, [1] by the high one
    DevBuf<char> ipc_stage;
    DevBuf<unsigned long long> seq;      // [0] ghost-update passes signalled, [1] mailbox all-reduces done: on the device, never reset
    // small all-reduces over peer memory (dfr_slab.cuh: SlabMail): every rank maps every rank's mailbox
    DevBuf<double> mbox;
    DevBuf<unsigned long long> mflag;
    SlabMail mail;
    bool mail_ok = false;
    std::vector<void *> ipc_more;
    // device-side particle exchange of replayed steps (dfr_slab.cuh: SlabXchg): my receive area + the neighbours', mapped
    DevBuf<char> xin;
    SlabXchg xchg;
    bool xchg_ok = false;
    bool h_stale = false;                // h_ranges / launch_nf are behind the device (replayed steps do not read back)
    long long sync_rows = 0;             // rows one ghost update moves (both directions), from the last exchange
    int64_t cap_syncs_static = 0, cap_syncs_div = 0, cap_syncs_prs = 0;  // ghost updates per replayed step / iteration
    int64_t *cap_sync_counter = nullptr;
  } slab;

  // SM-local scheduling of the gather kernels (dfr_kernels.cuh: VSched)
  DevBuf<unsigned int> sched_ctr;
  int sched_parity = 0, nsm = 1;
  std::map<const void *, int> resident_blocks;

  // bookkeeping
  int spec_div = 1, spec_prs = 2;
  int div_pred_streak = 2;  // consecutive steps whose divergence-iteration count matched the speculated one
  int fresh_steps = 4;      // steps since finalize / reset / load during which the list capacities are checked eagerly
  double device_ms = 0.0;
  int64_t launches = 0;
  int launch_nf = 0;

  // optional per-kernel timing (dfr_set_profiling): event pairs around every launch on the context's stream
  struct ProfPending {
    const char *name;
    cudaEvent_t e0, e1;
    int solver;  // -1 none, 0 divergence, 1 pressure
    int iter;    // iteration index of a speculatively enqueued solver kernel
  };
  struct ProfRow {
    double ms = 0.0;
    int64_t n = 0;
  };
  bool profiling = false;
  std::vector<ProfPending> prof_pending;
  std::vector<cudaEvent_t> prof_pool;
  std::map<std::string, ProfRow> prof_rows;
  int prof_solver = -1, prof_iter = -1;
};

namespace {

int fail(dfr_context *c, int code, const std::string &msg) {
  if (c) c->err = msg;
  return code;
}
#define CU(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      return fail(c, DFR_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));                 \
  } while (0)

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
inline int getenv_int(const char *name) {
  const char *v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : 0;
}
cudaEvent_t prof_event(dfr_context *c) {
  if (!c->prof_pool.empty()) {
    cudaEvent_t e = c->prof_pool.back();
    c->prof_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
void prof_begin(dfr_context *c, const char *name) {
  dfr_context::ProfPending p;
  p.name = name;
  p.e0 = prof_event(c);
  p.e1 = prof_event(c);
  p.solver = c->prof_solver;
  p.iter = c->prof_iter;
  cudaEventRecord(p.e0, c->stream);
  c->prof_pending.push_back(p);
}
void prof_end(dfr_context *c) { cudaEventRecord(c->prof_pending.back().e1, c->stream); }
// Resolve the pending event pairs (stream must be idle). Solver kernels enqueued speculatively for an
// iteration the device-side convergence test had already closed are booked under "<name> (idle)".
void prof_resolve(dfr_context *c, int used_div, int used_prs) {
  for (auto &p : c->prof_pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) {
      std::string key = p.name;
      if (p.solver >= 0 && p.iter >= (p.solver ? used_prs : used_div)) key += " (idle)";
      auto &row = c->prof_rows[key];
      row.ms += ms;
      row.n += 1;
    }
    c->prof_pool.push_back(p.e0);
    c->prof_pool.push_back(p.e1);
  }
  c->prof_pending.clear();
}
#define LAUNCH(c, kernel, grid, block, ...)                        \
  do {                                                             \
    if ((grid) > 0) {                                              \
      if ((c)->profiling) prof_begin((c), #kernel);                \
      kernel<<<(grid), (block), 0, (c)->ls>>>(__VA_ARGS__);        \
      if ((c)->profiling) prof_end((c));                           \
      if ((c)->capturing) (*(c)->cap_counter)++;                   \
      else (c)->launches++;                                        \
    }                                                              \
  } while (0)

// persistent launch of a gather kernel: SMs x resident blocks CTAs walk `nvb` virtual 128-particle blocks
VSched next_sched(dfr_context *c, int nvb) {
  VSched S;
  S.ctr = c->sched_ctr.p;
  S.parity = c->sched_parity;
  c->sched_parity ^= 1;
  S.nsm = c->nsm;
  S.nvb = nvb;
  S.per = (nvb + c->nsm - 1) / c->nsm;
  return S;
}
template <class K>
int persistent_grid(dfr_context *c, K kernel, int nvb) {
#if DFR_SM_LOCAL
  auto it = c->resident_blocks.find((const void *)kernel);
  if (it == c->resident_blocks.end()) {
    int nb = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, 128, 0) != cudaSuccess || nb < 1) nb = 1;
    it = c->resident_blocks.emplace((const void *)kernel, nb).first;
  }
  return std::max(1, std::min(nvb, c->nsm * it->second));
#else
  (void)kernel;
  return nvb;
#endif
}
#define PLAUNCH(c, kernel, nvb, ...)                                                          \
  do {                                                                                        \
    if ((nvb) > 0) {                                                                          \
      const int pg__ = persistent_grid((c), kernel, (nvb));                                   \
      if ((c)->profiling) prof_begin((c), #kernel);                                           \
      kernel<<<(pg__ + DFR_CTA_VB - 1) / DFR_CTA_VB, DFR_CTA_THREADS, 0, (c)->ls>>>(__VA_ARGS__, next_sched((c), (nvb))); \
      if ((c)->profiling) prof_end((c));                                                      \
      if ((c)->capturing) (*(c)->cap_counter)++;                                              \
      else (c)->launches++;                                                                   \
    }                                                                                         \
  } while (0)

int scan_u32(dfr_context *c, unsigned int *data, size_t n, unsigned int *total) {
  const int ntiles = cdiv((int64_t)n, SCAN_TILE);
  if ((size_t)ntiles > c->tile_sums.n) return fail(c, DFR_ERR_CAPACITY, "scan scratch too small");
  unsigned int *ts = c->on_side_branch ? c->tile_sums_side.p : c->tile_sums.p;  // two scans may be in flight in a recorded step
  LAUNCH(c, k_scan_tiles, ntiles, SCAN_THREADS, data, data, ts, n);
  LAUNCH(c, k_scan_sums, 1, 1024, ts, ntiles, total);
  if (ntiles > 1) LAUNCH(c, k_scan_add, cdiv((int64_t)n, 1024), 256, data, ts, n);
  return DFR_OK;
}

void host_body_record(dfr_context *c, HostBody &hb, int p_begin_dev) {
  // Dynamic3dRigidBody::determineMassProperties (Dynamic3dRigidBody.h:184-198)
  BodyDev &B = hb.dev0;
  std::memset(&B, 0, sizeof(B));
  const double r = c->cfg.particle_radius;
  const double volume = (4.0 / 3.0 * M_PI) * r * r * r;
  const double dm = volume * hb.density;
  double mass = 0.0;
  double I0[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int64_t i = 0; i < hb.n; i++) {
    mass += dm;
    const double x = hb.x_local[3 * i], y = hb.x_local[3 * i + 1], z = hb.x_local[3 * i + 2];
    const double rr = x * x + y * y + z * z;
    const double v[3] = {x, y, z};
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) I0[3 * a + b] += dm * ((a == b ? rr : 0.0) - v[a] * v[b]);
  }
  B.dynamic = hb.dynamic;
  B.animated = 0;
  B.p_begin = p_begin_dev;
  B.p_count = (int)hb.n;
  B.blk_begin = hb.blk_begin;
  B.blk_count = hb.blk_count;
  B.mass = mass;
  B.inv_mass = 1.0 / mass;
  for (int k = 0; k < 9; k++) B.I0.a[k] = I0[k];
  B.pos = B.pos0 = mk3(hb.pos[0], hb.pos[1], hb.pos[2]);
  B.q.w = hb.q[0]; B.q.x = hb.q[1]; B.q.y = hb.q[2]; B.q.z = hb.q[3];
  B.q0 = B.q;
  const m33 R = qrot(B.q);
  B.I = R * B.I0 * transpose(R);
  B.Iinv = inverse(B.I);
  B.vel = B.omega = mk3(0, 0, 0);
  B.init_v = mk3(hb.init_v[0], hb.init_v[1], hb.init_v[2]);
  B.init_omega = mk3(hb.init_w[0], hb.init_w[1], hb.init_w[2]);
  // BoundaryModel_Akinci2012::reset (:141-162)
  B.v_v0 = m33::identity();
  B.w_w0 = m33::identity();
}

GridView grid_fluid(dfr_context *c) {
  GridView g;
  g.cell_start = c->cell_start_f.p;
  g.sorted_src = nullptr;
  g.pos = c->pos[c->cur].p;
  g.base = 0;
  return g;
}
GridView grid_static(dfr_context *c) {
  GridView g;
  g.cell_start = c->cell_start_s.p;
  g.sorted_src = nullptr;
  g.pos = c->bpos.p;
  g.base = 0;
  return g;
}
GridView grid_dyn(dfr_context *c) {
  GridView g;
  g.cell_start = c->cell_start_d.p;
  g.sorted_src = c->sorted_src_d.p;
  g.pos = c->bpos.p + c->dyn_begin;
  g.base = c->dyn_begin;
  return g;
}
NbrList list_f(dfr_context *c) {
  NbrList l;
  l.cnt = c->cnt_f.p;
  l.idx = c->idx_f.p;
  l.cap = c->cap_f;
  return l;
}
NbrList list_b(dfr_context *c) {
  NbrList l;
  l.cnt = c->cnt_b.p;
  l.idx = c->idx_b.p;
  l.cap = c->cap_b;
  return l;
}

// cell table of the dynamic boundary particles (rebuilt whenever they moved)
int build_dyn_grid(dfr_context *c) {
  if (c->n_dyn_p == 0) return DFR_OK;
  const int nc = c->P.grid.ncells;
  cudaMemsetAsync(c->cell_start_d.p, 0, sizeof(unsigned int) * (nc + 1), c->ls);
  LAUNCH(c, k_bin_count, cdiv(c->n_dyn_p, 128), 128, c->P, c->bpos.p + c->dyn_begin, (const int *)nullptr, c->n_dyn_p,
         c->cell_start_d.p, c->cell_of_b.p, c->rank_b.p);
  int rc = scan_u32(c, c->cell_start_d.p, (size_t)nc + 1, nullptr);
  if (rc) return rc;
  LAUNCH(c, k_bin_scatter, cdiv(c->n_dyn_p, 128), 128, (const int *)nullptr, c->n_dyn_p, c->cell_start_d.p, c->cell_of_b.p,
         c->rank_b.p, c->sorted_src_d.p);
  LAUNCH(c, k_bin_sort_cells_by_particle, cdiv(c->n_dyn_p, 128), 128, (const int *)nullptr, c->n_dyn_p, c->cell_start_d.p,
         c->cell_of_b.p, c->rank_b.p, c->sorted_src_d.p);
  // cells from which a dynamic boundary particle may be within reach: the list build skips the 25-row walk over this
  // set for every fluid particle elsewhere (most of the fluid)
  cudaMemsetAsync(c->near_d.p, 0, (size_t)nc, c->ls);
  LAUNCH(c, k_mark_near_warp, cdiv((int64_t)c->n_dyn_p * 32, 128), 128, c->P, c->bpos.p + c->dyn_begin, c->n_dyn_p, c->near_d.p);
  return DFR_OK;
}

int build_neighbor_lists(dfr_context *c, bool dyn_grid_done = false);
int slab_exchange_and_sort(dfr_context *c);
int slab_exchange_device(dfr_context *c);
// CompactNSearch replacement: counting sort of the fluid into cell order + neighbour lists
int build_neighbors(dfr_context *c) {
  const int nc = c->P.grid.ncells;
  const int n = c->launch_nf;
  const int *nf_ptr = &c->dSt.p->nf;
  if (c->slab.on) {
    int rc = c->capturing ? slab_exchange_device(c) : slab_exchange_and_sort(c);
    if (rc) return rc;
    return build_neighbor_lists(c);
  }
  const int a = c->cur, b = 1 - c->cur;
  // Recorded steps: the grid of the dynamic boundary particles (8 small launches, independent of the fluid's sort) goes on
  // a side branch of the graph - fork here, join in front of the list build (DFR_NO_FORK=1: one chain as on the stream).
  const bool fork = c->capturing && c->n_dyn_p > 0 && getenv_int("DFR_NO_FORK") == 0;
  cudaStream_t main_stream = c->ls;
  if (fork) {
    if (!c->side_stream) {
      CU(cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking));
      CU(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    }
    CU(cudaEventRecord(c->ev_fork, main_stream));
    CU(cudaStreamWaitEvent(c->side_stream, c->ev_fork, 0));
    c->ls = c->side_stream;
    c->on_side_branch = true;
    int rcd = build_dyn_grid(c);
    c->on_side_branch = false;
    c->ls = main_stream;
    if (rcd) return rcd;
    CU(cudaEventRecord(c->ev_join, c->side_stream));
  }
  cudaMemsetAsync(c->cell_start_f.p, 0, sizeof(unsigned int) * (nc + 1), c->ls);
  LAUNCH(c, k_bin_count, cdiv(n, 128), 128, c->P, c->pos[a].p, nf_ptr, 0, c->cell_start_f.p, c->cell_of_p.p, c->rank_in_cell.p);
  int rc = scan_u32(c, c->cell_start_f.p, (size_t)nc + 1, nullptr);
  if (rc) return rc;
  LAUNCH(c, k_bin_scatter, cdiv(n, 128), 128, nf_ptr, 0, c->cell_start_f.p, c->cell_of_p.p, c->rank_in_cell.p, c->sorted_src_f.p);
  LAUNCH(c, k_bin_sort_cells_by_particle, cdiv(n, 128), 128, nf_ptr, 0, c->cell_start_f.p, c->cell_of_p.p, c->rank_in_cell.p,
         c->sorted_src_f.p);
  LAUNCH(c, k_permute_fluid, cdiv(n, 128), 128, c->dSt.p, c->sorted_src_f.p, c->pos[a].p, c->vel[c->vcur].p, c->kappa[a].p,
         c->kappav[a].p, c->pid[a].p, c->pstate[a].p, c->pos[b].p, c->vel[1 - c->vcur].p, c->kappa[b].p, c->kappav[b].p,
         c->pid[b].p, c->pstate[b].p);
  c->cur = b;
  c->vcur = 1 - c->vcur;
  if (fork) CU(cudaStreamWaitEvent(c->ls, c->ev_join, 0));
  return build_neighbor_lists(c, fork);
}

// neighbour lists of the sorted fluid (and the dynamic-boundary -> fluid rows)
int build_neighbor_lists(dfr_context *c, bool dyn_grid_done) {
  const int n = c->launch_nf;
  int rc = dyn_grid_done ? DFR_OK : build_dyn_grid(c);
  if (rc) return rc;
  // Recorded steps: the rows of the dynamic boundary particles (count, scan, fill: they only read the sorted fluid) are
  // built on a side branch of the graph while k_nbr_build runs (dyn_grid_done says build_neighbors set the branch up).
  const bool fork = dyn_grid_done && c->capturing && c->n_dyn_p > 0 && c->side_stream != nullptr;
  cudaStream_t main_stream = c->ls;
  if (fork) {
    CU(cudaEventRecord(c->ev_fork, main_stream));
    CU(cudaStreamWaitEvent(c->side_stream, c->ev_fork, 0));
    c->ls = c->side_stream;
    c->on_side_branch = true;
  } else
    PLAUNCH(c, k_nbr_build, cdiv(n, 128), c->P, c->dSt.p, c->pos[c->cur].p, grid_fluid(c), grid_static(c), grid_dyn(c),
           c->n_static_p > 0 ? 1 : 0, c->n_dyn_p > 0 ? 1 : 0, c->cnt_f.p, c->cnt_b.p, c->idx_f.p, c->idx_b.p, c->cap_f, c->cap_b,
           c->near_s.p, c->near_d.p);
  if (c->n_dyn_p > 0) {
    cudaMemsetAsync(c->off_d.p, 0, sizeof(unsigned int) * (c->n_dyn_p + 1), c->ls);
    LAUNCH(c, k_dnbr_count, cdiv((int64_t)c->n_dyn_p * 32, 128), 128, c->P, c->bpos.p, c->dyn_begin, c->n_dyn_p, grid_fluid(c), c->off_d.p);
    rc = scan_u32(c, c->off_d.p, (size_t)c->n_dyn_p + 1, nullptr);
    if (rc) return rc;
    LAUNCH(c, k_dnbr_fill, cdiv((int64_t)c->n_dyn_p * 32, 128), 128, c->P, c->dSt.p, c->bpos.p, c->dyn_begin, c->n_dyn_p, grid_fluid(c),
           c->off_d.p, c->idx_d.p, c->cap_d);
  }
  if (fork) {
    c->on_side_branch = false;
    c->ls = main_stream;
    CU(cudaEventRecord(c->ev_join, c->side_stream));
    PLAUNCH(c, k_nbr_build, cdiv(n, 128), c->P, c->dSt.p, c->pos[c->cur].p, grid_fluid(c), grid_static(c), grid_dyn(c),
           c->n_static_p > 0 ? 1 : 0, c->n_dyn_p > 0 ? 1 : 0, c->cnt_f.p, c->cnt_b.p, c->idx_f.p, c->idx_b.p, c->cap_f, c->cap_b,
           c->near_s.p, c->near_d.p);
    CU(cudaStreamWaitEvent(c->ls, c->ev_join, 0));
  }
  return DFR_OK;
}

int compute_boundary_volumes(dfr_context *c) {
  if (c->n_b == 0) return DFR_OK;
  LAUNCH(c, k_boundary_volume, cdiv(c->n_b, 128), 128, c->P, c->bpos.p, c->n_b, c->n_static_p, grid_static(c), grid_dyn(c),
         c->n_static_p > 0 ? 1 : 0, c->n_dyn_p > 0 ? 1 : 0, c->bvol.p);
  LAUNCH(c, k_store_volume, cdiv(c->n_b, 128), 128, c->bpos.p, c->bvol.p, c->n_b);
  return DFR_OK;
}

int sync_state(dfr_context *c) {
  CU(cudaMemcpyAsync(c->hSt, c->dSt.p, sizeof(StepState), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (c->hSt->error_flags & 32) return fail(c, DFR_ERR_STATE, "slab: timed out waiting for a neighbour's ghost rows (did a peer fail?)");
  if (c->hSt->error_flags & 64) return fail(c, DFR_ERR_STATE, "slab: ghost layers do not match the neighbours' boundary layers");
  if (c->hSt->error_flags & 16) return fail(c, DFR_ERR_STATE, "slab: a particle moved further than one ghost layer in one step");
  if (c->hSt->error_flags & 8) return fail(c, DFR_ERR_CAPACITY, "slab: export buffer too small");
  if (c->hSt->error_flags) {
    char buf[160];
    std::snprintf(buf, sizeof(buf), "neighbour list capacity exceeded (flags %d; longest rows f=%u b=%u, d entries=%u; cap f=%d b=%d d=%u)",
                  c->hSt->error_flags, c->hSt->list_used_f, c->hSt->list_used_b, c->hSt->list_used_d, c->cap_f, c->cap_b, c->cap_d);
    return fail(c, DFR_ERR_CAPACITY, buf);
  }
  return DFR_OK;
}

// ---- CUDA-graph stepping ---------------------------------------------------------------------------------------
// A step is recorded once per buffer parity (the per-step re-sort flips pos/kappa/id buffers; the velocity buffer
// flips twice per step) by stream capture of launch_step() and replayed with one cudaGraphLaunch.  The two Jacobi
// loops are conditional WHILE nodes (condition written by k_residual_finish: TimeStepDiffDFSPH.cpp:711-743, 828-861),
// and the whole step hangs in an IF node opened by k_step_gate, so that dfr_run_trajectory can enqueue steps in batches
// without looking at `finished` after every one.  Nothing in a replayed step needs the host: no speculation, no
// read-back between the solves.
// Not recorded (the stream path below stays): slab-decomposed contexts on the NCCL transport (collective calls and
// host-side exchange sizes; with the peer-memory transport a slab step is recorded like any other, slab_exchange_device),
// per-kernel profiling, steps in which the list capacities are being watched, DFR_NO_GRAPH=1.
#define CUG(call)                                                                                        \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess) {                                                                            \
      cudaGetLastError();                                                                                \
      return fail(c, DFR_ERR_CUDA, std::string("step graph: ") + #call + ": " + cudaGetErrorString(e__)); \
    }                                                                                                    \
  } while (0)

int launch_step(dfr_context *c);

// A condition handle of the graph c->ls is recording into (value `default_value` at every launch of the graph).
int cap_make_handle(dfr_context *c, unsigned int default_value, unsigned long long *handle_out) {
  cudaStreamCaptureStatus status;
  cudaGraph_t g = nullptr;
  CUG(cudaStreamGetCaptureInfo_v2(c->ls, &status, nullptr, &g, nullptr, nullptr));
  if (status != cudaStreamCaptureStatusActive) return fail(c, DFR_ERR_STATE, "step graph: stream is not capturing");
  cudaGraphConditionalHandle h;
  CUG(cudaGraphConditionalHandleCreate(&h, g, default_value, cudaGraphCondAssignDefault));
  *handle_out = (unsigned long long)h;
  return DFR_OK;
}
// Adds a conditional node after everything captured so far on c->ls and redirects the launches to its body graph.
int cap_begin_conditional(dfr_context *c, cudaGraphConditionalNodeType type, unsigned long long handle) {
  if (c->cap_depth + 1 >= 3) return fail(c, DFR_ERR_STATE, "step graph: conditional nodes nested too deep");
  cudaStreamCaptureStatus status;
  cudaGraph_t g = nullptr;
  const cudaGraphNode_t *deps = nullptr;
  size_t ndeps = 0;
  CUG(cudaStreamGetCaptureInfo_v2(c->ls, &status, nullptr, &g, &deps, &ndeps));
  if (status != cudaStreamCaptureStatusActive) return fail(c, DFR_ERR_STATE, "step graph: stream is not capturing");
  const cudaGraphConditionalHandle h = (cudaGraphConditionalHandle)handle;
  cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
  np.conditional.handle = h;
  np.conditional.type = type;
  np.conditional.size = 1;
  cudaGraphNode_t node;
  CUG(cudaGraphAddNode(&node, g, deps, ndeps, &np));
  cudaGraph_t body = np.conditional.phGraph_out[0];
  CUG(cudaStreamUpdateCaptureDependencies(c->ls, &node, 1, cudaStreamSetCaptureDependencies));
  cudaStream_t bs = c->cap_stream[c->cap_depth + 1];
  CUG(cudaStreamBeginCaptureToGraph(bs, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
  c->cap_depth++;
  c->ls = bs;
  return DFR_OK;
}
int cap_end_conditional(dfr_context *c) {
  cudaGraph_t body = nullptr;
  CUG(cudaStreamEndCapture(c->ls, &body));
  c->cap_depth--;
  c->ls = c->cap_stream[c->cap_depth];
  return DFR_OK;
}
void drop_step_graphs(dfr_context *c) {
  for (auto &row : c->step_graph)
    for (auto &sg : row) {
      if (sg.exec) cudaGraphExecDestroy(sg.exec);
      if (sg.graph) cudaGraphDestroy(sg.graph);
      sg = dfr_context::StepGraph();
    }
}
bool graph_stepping_possible(const dfr_context *c) {
  if (c->slab.on && !(c->slab.p2p && c->slab.mail_ok)) return false;  // NCCL transport: collective calls inside the solves
  return !c->graph_broken && !c->profiling && getenv_int("DFR_NO_GRAPH") == 0 && getenv_int("DFR_NO_FUSION") != 3;
}
// Records the step that starts with buffer parity c->cur; `gated` = the step is skipped once the trajectory has finished.
// Host-side buffer indices are restored afterwards (recording executes nothing).
int capture_step_graph(dfr_context *c, int gated) {
  dfr_context::StepGraph &sg = c->step_graph[gated][c->cur];
  for (auto &cs : c->cap_stream)
    if (!cs) CUG(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
  const int cur0 = c->cur, vcur0 = c->vcur, parity0 = c->sched_parity, nf0 = c->launch_nf;
  sg = dfr_context::StepGraph();
  if (c->slab.on) {  // the number of local particles changes with every exchange: grids for the capacity, kernels test the range
    c->launch_nf = (int)c->nf_cap;
    c->slab.cap_syncs_static = c->slab.cap_syncs_div = c->slab.cap_syncs_prs = 0;
    c->slab.cap_sync_counter = &c->slab.cap_syncs_static;
  }
  c->capturing = true;
  c->cap_target = &sg;
  c->cap_counter = &sg.n_static;
  c->cap_depth = 0;
  c->ls = c->cap_stream[0];
  int rc = DFR_OK;
  cudaError_t e = cudaStreamBeginCapture(c->ls, cudaStreamCaptureModeRelaxed);
  if (e != cudaSuccess) rc = fail(c, DFR_ERR_CUDA, std::string("step graph: cudaStreamBeginCapture: ") + cudaGetErrorString(e));
  if (!rc && gated) {  // k_step_gate -> IF node holding the step
    unsigned long long gate = 0;
    rc = cap_make_handle(c, 1u, &gate);
    if (!rc) {
      k_step_gate<<<1, 1, 0, c->ls>>>(c->dSt.p, gate, 1);
      rc = cap_begin_conditional(c, cudaGraphCondTypeIf, gate);
    }
    if (!rc) {
      rc = launch_step(c);
      if (!rc) rc = cap_end_conditional(c);
    }
  } else if (!rc)
    rc = launch_step(c);
  cudaGraph_t g = nullptr;
  e = cudaStreamEndCapture(c->cap_stream[0], &g);
  c->capturing = false;
  c->on_side_branch = false;
  c->cap_target = nullptr;
  c->cap_counter = nullptr;
  c->ls = c->stream;
  c->cur = cur0;
  c->vcur = vcur0;
  c->sched_parity = parity0;
  c->launch_nf = nf0;
  if (!rc && e != cudaSuccess) rc = fail(c, DFR_ERR_CUDA, std::string("step graph: cudaStreamEndCapture: ") + cudaGetErrorString(e));
  if (rc) {
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
    return rc;
  }
  sg.graph = g;
  e = cudaGraphInstantiate(&sg.exec, g, 0);
  if (e != cudaSuccess) {
    cudaGetLastError();
    cudaGraphDestroy(g);
    sg = dfr_context::StepGraph();
    return fail(c, DFR_ERR_CUDA, std::string("step graph: cudaGraphInstantiate: ") + cudaGetErrorString(e));
  }
  return DFR_OK;
}

// ---- NCCL, loaded at run time (libnccl.so.2: the copy torch already mapped, else the system one) ----
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi *nccl_api(std::string *why) {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.handle) api.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) {
      bool ok = true;
      auto sym = [&](const char *name) {
        void *p = dlsym(api.handle, name);
        if (!p) ok = false;
        return p;
      };
      api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
      api.Send = (decltype(api.Send))sym("ncclSend");
      api.Recv = (decltype(api.Recv))sym("ncclRecv");
      api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
      api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
      api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
      api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
      if (!ok) {
        dlclose(api.handle);
        api.handle = nullptr;
      }
    }
  }
  if (!api.handle) {
    if (why) *why = "libnccl.so.2 could not be loaded (slab decomposition needs NCCL)";
    return nullptr;
  }
  return &api;
}
#define NC(call)                                                                                                   \
  do {                                                                                                             \
    ncclResult_t r__ = (call);                                                                                     \
    if (r__ != ncclSuccess) return fail(c, DFR_ERR_CUDA, std::string(#call) + ": " + nccl_api(nullptr)->GetErrorString(r__)); \
  } while (0)

enum { GA_XK = 0, GA_XRHO = 1, GA_NORMAL = 2, GA_VEL0 = 3, GA_VEL1 = 4 };
// where the producing kernel has to mirror its boundary-layer rows (nothing unless the peer-memory transport is on)
GhostOut ghost_out(dfr_context *c, int which) {
  GhostOut g;
  std::memset(&g, 0, sizeof(g));
  auto &S = c->slab;
  if (!S.on || !S.p2p) return g;
  // row ranges and the low neighbour's own_end are read on the device (StepState::slab_ranges; the header of my receive
  // area holds what the low neighbour sent with the layer counts of the last exchange)
  if (S.G.has_lo) g.lo = S.peer_lo[which];
  if (S.G.has_hi) g.hi = S.peer_hi[which];
  g.ranges = c->dSt.p->slab_ranges;
  g.lo_nb_own_end = xchg_hdr(S.xin.p) + XH_LO_OWN_END;
  return g;
}

// ghost update of one per-particle array (element size `esz` bytes): my boundary layers -> the neighbours' ghost layers
int slab_sync(dfr_context *c, void *buf, size_t esz) {
  auto &S = c->slab;
  if (S.p2p) {  // the rows were written by the producing kernel; tell the neighbours and wait for theirs
    LAUNCH(c, k_slab_signal_wait, 1, 32, S.G.has_lo ? S.peer_lo_flags + 1 : (unsigned long long *)nullptr,
           S.G.has_hi ? S.peer_hi_flags + 0 : (unsigned long long *)nullptr, (volatile unsigned long long *)S.flags.p, S.seq.p + 0,
           5000000000ull, &c->dSt.p->error_flags);
    if (c->capturing)
      (*S.cap_sync_counter)++;  // replayed steps: traffic is accounted from these counts after the step (account_graph_launches)
    else
      S.exchanged_bytes += S.sync_rows * (long long)esz;
    return DFR_OK;
  }
  NcclApi *N = nccl_api(nullptr);
  const int own_begin = S.h_ranges[0], own_end = S.h_ranges[1], bl_lo_end = S.h_ranges[2], bl_hi_begin = S.h_ranges[3], nf = S.h_ranges[4];
  char *b = (char *)buf;
  NC(N->GroupStart());
  if (S.G.has_lo) {
    NC(N->Send(b + (size_t)own_begin * esz, (size_t)(bl_lo_end - own_begin) * esz, ncclChar, S.rank - 1, S.comm, c->stream));
    NC(N->Recv(b, (size_t)own_begin * esz, ncclChar, S.rank - 1, S.comm, c->stream));
    S.exchanged_bytes += (long long)((bl_lo_end - own_begin) + own_begin) * (long long)esz;
  }
  if (S.G.has_hi) {
    NC(N->Send(b + (size_t)bl_hi_begin * esz, (size_t)(own_end - bl_hi_begin) * esz, ncclChar, S.rank + 1, S.comm, c->stream));
    NC(N->Recv(b + (size_t)own_end * esz, (size_t)(nf - own_end) * esz, ncclChar, S.rank + 1, S.comm, c->stream));
    S.exchanged_bytes += (long long)((own_end - bl_hi_begin) + (nf - own_end)) * (long long)esz;
  }
  NC(N->GroupEnd());
  c->launches++;
  return DFR_OK;
}
int slab_allreduce(dfr_context *c, void *buf, size_t count, ncclDataType_t type, ncclRedOp_t op) {
  NcclApi *N = nccl_api(nullptr);
  NC(N->AllReduce(buf, buf, count, type, op, c->slab.comm, c->stream));
  c->launches++;
  return DFR_OK;
}
#define SLAB_SYNC(c, buf, esz)                    \
  do {                                            \
    if ((c)->slab.on) {                           \
      int rc__ = slab_sync((c), (buf), (esz));    \
      if (rc__) return rc__;                      \
    }                                             \
  } while (0)

// Map the neighbours' gathered arrays and flag words into this process (cudaIpc) so that the producing kernels can
// store boundary rows straight into the neighbours' ghost ranges over NVLink.  All ranks agree (all-reduce) on whether
// the mapping worked; otherwise everybody stays on NCCL send/recv.  DFR_SLAB_TRANSPORT=nccl forces the fallback.
int slab_p2p_setup(dfr_context *c) {
  auto &S = c->slab;
  NcclApi *N = nccl_api(nullptr);
  const char *env = getenv("DFR_SLAB_TRANSPORT");
  int want = !(env && std::string(env) == "nccl");
  CU(S.flags.alloc(8));
  {
    // Receive area of the device-side particle exchange.  A sender addresses rows in the RECEIVER's layout, so every
    // rank uses the same row capacity: the largest send_cap of the job (the estimates differ from slab to slab; with
    // per-rank capacities the velocity / misc rows landed at the wrong offsets - found by the 4- and 8-rank slab checks).
    int cap = S.send_cap;
    CU(cudaMemcpyAsync(S.counts.p, &cap, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    NC(N->AllReduce(S.counts.p, S.counts.p, 1, ncclInt, ncclMax, S.comm, c->stream));
    CU(cudaMemcpyAsync(&cap, S.counts.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    S.xin.free();
    CU(S.xin.alloc(xchg_bytes(cap)));
    S.xchg.mine = S.xin.p;
    S.xchg.cap = cap;
  }
  const int NH = 7;
  CU(S.ipc_stage.alloc(3 * NH * sizeof(cudaIpcMemHandle_t) + 16));
  void *mine[NH] = {c->xk.p, c->xrho.p, c->normal.p, c->vel[0].p, c->vel[1].p, S.flags.p, S.xin.p};
  std::vector<cudaIpcMemHandle_t> h(3 * NH);
  std::memset(h.data(), 0, h.size() * sizeof(cudaIpcMemHandle_t));
  int ok = want;
  for (int k = 0; k < NH && ok; k++)
    if (cudaIpcGetMemHandle(&h[k], mine[k]) != cudaSuccess) {
      cudaGetLastError();
      ok = 0;
    }
  const size_t hb = NH * sizeof(cudaIpcMemHandle_t);
  CU(cudaMemcpyAsync(S.ipc_stage.p, h.data(), hb, cudaMemcpyHostToDevice, c->stream));
  NC(N->GroupStart());
  if (S.G.has_lo) {
    NC(N->Send(S.ipc_stage.p, hb, ncclChar, S.rank - 1, S.comm, c->stream));
    NC(N->Recv(S.ipc_stage.p + hb, hb, ncclChar, S.rank - 1, S.comm, c->stream));
  }
  if (S.G.has_hi) {
    NC(N->Send(S.ipc_stage.p, hb, ncclChar, S.rank + 1, S.comm, c->stream));
    NC(N->Recv(S.ipc_stage.p + 2 * hb, hb, ncclChar, S.rank + 1, S.comm, c->stream));
  }
  NC(N->GroupEnd());
  CU(cudaMemcpyAsync(h.data(), S.ipc_stage.p, 3 * hb, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  for (int side = 0; side < 2 && ok; side++) {
    if (!(side == 0 ? S.G.has_lo : S.G.has_hi)) continue;
    for (int k = 0; k < NH && ok; k++) {
      void *ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, h[(1 + side) * NH + k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
        break;
      }
      S.ipc_opened[side * NH + k] = ptr;
      if (k < 5)
        (side == 0 ? S.peer_lo : S.peer_hi)[k] = (double4 *)ptr;
      else if (k == 5)
        (side == 0 ? S.peer_lo_flags : S.peer_hi_flags) = (unsigned long long *)ptr;
      else
        (side == 0 ? S.xchg.peer_lo : S.xchg.peer_hi) = (char *)ptr;
    }
  }
  // ---- the mailboxes of ALL ranks: residual sums, the CFL maximum and the per-body rows are all-reduced by one small
  // kernel over peer memory (dfr_slab.cuh: mailbox_allreduce) instead of a collective launch ----
  CU(S.seq.alloc(2));
  const int mcap = std::max<int>(8, (int)std::max<size_t>(c->bodies.size(), 1) * ACC_N);
  CU(S.mbox.alloc((size_t)2 * S.n * mcap));
  CU(S.mflag.alloc((size_t)S.n));
  int mail = (ok && S.n <= DFR_MAX_SLABS) ? 1 : 0;
  {
    const size_t hsz = sizeof(cudaIpcMemHandle_t), per = 2 * hsz;
    std::vector<cudaIpcMemHandle_t> mh(2 * (size_t)(S.n + 1));
    std::memset(mh.data(), 0, mh.size() * hsz);
    if (mail && (cudaIpcGetMemHandle(&mh[0], S.mbox.p) != cudaSuccess || cudaIpcGetMemHandle(&mh[1], S.mflag.p) != cudaSuccess)) {
      cudaGetLastError();
      mail = 0;
    }
    DevBuf<char> mstage;
    CU(mstage.alloc(per * (size_t)(S.n + 1)));
    CU(cudaMemcpyAsync(mstage.p, mh.data(), per, cudaMemcpyHostToDevice, c->stream));
    NC(N->GroupStart());
    for (int p = 0; p < S.n; p++) {
      if (p == S.rank) continue;
      NC(N->Send(mstage.p, per, ncclChar, p, S.comm, c->stream));
      NC(N->Recv(mstage.p + per * (size_t)(1 + p), per, ncclChar, p, S.comm, c->stream));
    }
    NC(N->GroupEnd());
    CU(cudaMemcpyAsync(mh.data(), mstage.p, per * (size_t)(S.n + 1), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    mstage.free();
    std::memset(&S.mail, 0, sizeof(S.mail));
    S.mail.n = S.n;
    S.mail.rank = S.rank;
    S.mail.cap = mcap;
    S.mail.seq = S.seq.p + 1;
    for (int p = 0; p < S.n && mail; p++) {
      if (p == S.rank) {
        S.mail.box[p] = S.mbox.p;
        S.mail.flag[p] = S.mflag.p;
        continue;
      }
      void *pb = nullptr, *pf = nullptr;
      if (cudaIpcOpenMemHandle(&pb, mh[2 * (size_t)(1 + p)], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
          cudaIpcOpenMemHandle(&pf, mh[2 * (size_t)(1 + p) + 1], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        mail = 0;
        if (pb) S.ipc_more.push_back(pb);
        break;
      }
      S.ipc_more.push_back(pb);
      S.ipc_more.push_back(pf);
      S.mail.box[p] = (double *)pb;
      S.mail.flag[p] = (unsigned long long *)pf;
    }
  }
  if (getenv_int("DFR_SLAB_NO_MAILBOX")) mail = 0;  // A/B: keep the NCCL all-reduces
  // everybody or nobody
  int *flag = (int *)(S.ipc_stage.p + 3 * hb);
  int both[2] = {ok, mail};
  CU(cudaMemcpyAsync(flag, both, 2 * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  NC(N->AllReduce(flag, flag, 2, ncclInt, ncclMin, S.comm, c->stream));
  CU(cudaMemcpyAsync(both, flag, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  S.p2p = both[0] != 0;
  S.mail_ok = S.p2p && both[1] != 0;
  // the particle exchange of replayed steps runs over the same mappings (DFR_SLAB_HOST_EXCHANGE=1: keep the NCCL exchange
  // with its two read-backs at the head of every step; the variable must be set on all ranks alike)
  S.xchg_ok = S.mail_ok && getenv_int("DFR_SLAB_HOST_EXCHANGE") == 0;
  return DFR_OK;
}

// Step start in slab mode: export the particles that now sit in my boundary layers (or beyond: emigrants), import the
// neighbours', and re-sort.  Replaces the head of build_neighbors.
int slab_exchange_and_sort(dfr_context *c) {
  auto &S = c->slab;
  NcclApi *N = nccl_api(nullptr);
  const int a = c->cur, b = 1 - c->cur;
  if (S.h_stale) {  // replayed steps went by: every dfr_step ends with a read-back of the step state
    std::memcpy(S.h_ranges, c->hSt->slab_ranges, sizeof(S.h_ranges));
    S.h_stale = false;
  }
  const int own_begin = S.h_ranges[0], own_end = S.h_ranges[1];
  const int n_own = own_end - own_begin;
  int *const hdr = xchg_hdr(S.xin.p);
  CU(cudaMemsetAsync(S.counts.p, 0, 8 * sizeof(int), c->stream));
  LAUNCH(c, k_slab_select, cdiv(n_own, 128), 128, c->P, S.G, c->dSt.p, c->pos[a].p, c->vel[c->vcur].p, c->kappa[a].p, c->kappav[a].p,
         c->pid[a].p, c->pstate[a].p, S.send_cap, S.s_pos[0].p, S.s_vel[0].p, S.s_misc[0].p, S.s_pos[1].p, S.s_vel[1].p, S.s_misc[1].p,
         S.counts.p, &c->dSt.p->error_flags);
  // counts: mine to the host, and to the neighbours
  NC(N->GroupStart());
  if (S.G.has_lo) {
    NC(N->Send(S.counts.p + 0, 1, ncclInt, S.rank - 1, S.comm, c->stream));
    NC(N->Recv(S.counts.p + 2, 1, ncclInt, S.rank - 1, S.comm, c->stream));
  }
  if (S.G.has_hi) {
    NC(N->Send(S.counts.p + 1, 1, ncclInt, S.rank + 1, S.comm, c->stream));
    NC(N->Recv(S.counts.p + 3, 1, ncclInt, S.rank + 1, S.comm, c->stream));
  }
  NC(N->GroupEnd());
  CU(cudaMemcpyAsync(S.h_counts, S.counts.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  const int send_lo = S.G.has_lo ? S.h_counts[0] : 0, send_hi = S.G.has_hi ? S.h_counts[1] : 0;
  const int recv_lo = S.G.has_lo ? S.h_counts[2] : 0, recv_hi = S.G.has_hi ? S.h_counts[3] : 0;
  if (send_lo > S.send_cap || send_hi > S.send_cap || recv_lo > S.send_cap || recv_hi > S.send_cap)
    return fail(c, DFR_ERR_CAPACITY, "slab exchange buffer too small");
  if ((int64_t)own_end + recv_lo + recv_hi > c->nf_cap) return fail(c, DFR_ERR_CAPACITY, "slab: local particle capacity exceeded");
  // payload: positions and velocities straight behind my owned range, (kappa, kappa_v, id, state) through a scratch buffer
  double4 *pos = c->pos[a].p, *vel = c->vel[c->vcur].p;
  NC(N->GroupStart());
  if (S.G.has_lo) {
    NC(N->Send(S.s_pos[0].p, (size_t)send_lo * 4, ncclDouble, S.rank - 1, S.comm, c->stream));
    NC(N->Send(S.s_vel[0].p, (size_t)send_lo * 4, ncclDouble, S.rank - 1, S.comm, c->stream));
    NC(N->Send(S.s_misc[0].p, (size_t)send_lo * 4, ncclDouble, S.rank - 1, S.comm, c->stream));
    NC(N->Recv(pos + own_end, (size_t)recv_lo * 4, ncclDouble, S.rank - 1, S.comm, c->stream));
    NC(N->Recv(vel + own_end, (size_t)recv_lo * 4, ncclDouble, S.rank - 1, S.comm, c->stream));
    NC(N->Recv(S.r_misc.p, (size_t)recv_lo * 4, ncclDouble, S.rank - 1, S.comm, c->stream));
  }
  if (S.G.has_hi) {
    NC(N->Send(S.s_pos[1].p, (size_t)send_hi * 4, ncclDouble, S.rank + 1, S.comm, c->stream));
    NC(N->Send(S.s_vel[1].p, (size_t)send_hi * 4, ncclDouble, S.rank + 1, S.comm, c->stream));
    NC(N->Send(S.s_misc[1].p, (size_t)send_hi * 4, ncclDouble, S.rank + 1, S.comm, c->stream));
    NC(N->Recv(pos + own_end + recv_lo, (size_t)recv_hi * 4, ncclDouble, S.rank + 1, S.comm, c->stream));
    NC(N->Recv(vel + own_end + recv_lo, (size_t)recv_hi * 4, ncclDouble, S.rank + 1, S.comm, c->stream));
    NC(N->Recv(S.r_misc.p + recv_lo, (size_t)recv_hi * 4, ncclDouble, S.rank + 1, S.comm, c->stream));
  }
  NC(N->GroupEnd());
  c->launches++;
  S.exchanged_bytes += (long long)(send_lo + send_hi + recv_lo + recv_hi) * 96;
  const int n_recv = recv_lo + recv_hi;
  LAUNCH(c, k_slab_unpack, cdiv(n_recv, 128), 128, S.r_misc.p, n_recv, own_end, c->kappa[a].p, c->kappav[a].p, c->pid[a].p, c->pstate[a].p);
  // sort [own_begin, own_end + n_recv) by (cell, id) into the other buffers
  const int n_src = n_own + n_recv;
  const int nc = c->P.grid.ncells;
  c->launch_nf = n_src;
  LAUNCH(c, k_slab_set_nf, 1, 32, c->dSt.p, n_src);
  CU(cudaMemsetAsync(c->cell_start_f.p, 0, sizeof(unsigned int) * (nc + 1), c->stream));
  LAUNCH(c, k_bin_count, cdiv(n_src, 128), 128, c->P, pos + own_begin, (const int *)nullptr, n_src, c->cell_start_f.p, c->cell_of_p.p,
         c->rank_in_cell.p);
  int rc = scan_u32(c, c->cell_start_f.p, (size_t)nc + 1, nullptr);
  if (rc) return rc;
  LAUNCH(c, k_bin_scatter, cdiv(n_src, 128), 128, (const int *)nullptr, n_src, c->cell_start_f.p, c->cell_of_p.p, c->rank_in_cell.p,
         c->sorted_src_f.p);
  LAUNCH(c, k_bin_sort_cells_by_id, cdiv(n_src, 128), 128, (const int *)nullptr, n_src, (const int *)nullptr, c->cell_start_f.p, c->cell_of_p.p,
         c->rank_in_cell.p, c->sorted_src_f.p, c->pid[a].p + own_begin);
  LAUNCH(c, k_permute_fluid, cdiv(n_src, 128), 128, c->dSt.p, c->sorted_src_f.p, pos + own_begin, vel + own_begin, c->kappa[a].p + own_begin,
         c->kappav[a].p + own_begin, c->pid[a].p + own_begin, c->pstate[a].p + own_begin, c->pos[b].p, c->vel[1 - c->vcur].p, c->kappa[b].p,
         c->kappav[b].p, c->pid[b].p, c->pstate[b].p);
  c->cur = b;
  c->vcur = 1 - c->vcur;
  LAUNCH(c, k_slab_ranges, 1, 32, c->P, S.G, c->dSt.p, c->cell_start_f.p);
  // both sides of a plane must agree on the shared layers before any ghost update is posted (a mismatched send/recv
  // pair would hang): tell the neighbours how many particles my boundary layers hold, compare with my ghost layers
  NC(N->GroupStart());
  if (S.G.has_lo) {
    NC(N->Send(&c->dSt.p->slab_ranges[5], 1, ncclInt, S.rank - 1, S.comm, c->stream));
    NC(N->Recv(hdr + XH_LO_BL, 2, ncclInt, S.rank - 1, S.comm, c->stream));  // its bl_hi count, its own_end
  }
  if (S.G.has_hi) {
    NC(N->Send(&c->dSt.p->slab_ranges[6], 2, ncclInt, S.rank + 1, S.comm, c->stream));  // my bl_hi count, my own_end (ranges[7])
    NC(N->Recv(hdr + XH_HI_BL, 1, ncclInt, S.rank + 1, S.comm, c->stream));
  }
  NC(N->GroupEnd());
  CU(cudaMemcpyAsync(S.h_counts, hdr, 8 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  rc = sync_state(c);
  if (rc) return rc;
  std::memcpy(S.h_ranges, c->hSt->slab_ranges, sizeof(S.h_ranges));
  const int ghost_lo = S.h_ranges[0], ghost_hi = S.h_ranges[4] - S.h_ranges[1];
  if ((S.G.has_lo && S.h_counts[2] != ghost_lo) || (S.G.has_hi && S.h_counts[4] != ghost_hi)) {
    char buf[200];
    std::snprintf(buf, sizeof(buf), "slab %d: ghost layers (%d, %d) do not match the neighbours' boundary layers (%d, %d)", S.rank, ghost_lo,
                  ghost_hi, S.G.has_lo ? S.h_counts[2] : 0, S.G.has_hi ? S.h_counts[4] : 0);
    return fail(c, DFR_ERR_STATE, buf);
  }
  S.sync_rows = (S.G.has_lo ? (S.h_ranges[2] - S.h_ranges[0]) + S.h_ranges[0] : 0) +
                (S.G.has_hi ? (S.h_ranges[1] - S.h_ranges[3]) + (S.h_ranges[4] - S.h_ranges[1]) : 0);
  return DFR_OK;
}

// The same on the device (replayed steps, dfr_slab.cuh "Device-side particle exchange"): peer stores + two flag passes;
// ranges, counts and capacities are read and checked by the kernels, grids are sized for the capacities.
int slab_exchange_device(dfr_context *c) {
  auto &S = c->slab;
  const int a = c->cur, b = 1 - c->cur;
  const int gcap = cdiv(c->nf_cap, 128);
  const int nc = c->P.grid.ncells;
  const SlabXchg &X = S.xchg;
  StepState *st = c->dSt.p;
  unsigned long long *flo = S.G.has_lo ? S.peer_lo_flags + 1 : (unsigned long long *)nullptr;
  unsigned long long *fhi = S.G.has_hi ? S.peer_hi_flags + 0 : (unsigned long long *)nullptr;
  // I am the high neighbour of my low neighbour: my rows go into the "from high" half of its area, and vice versa
  double4 *nul = nullptr;
  CU(cudaMemsetAsync(S.counts.p, 0, 8 * sizeof(int), c->ls));
  LAUNCH(c, k_slab_select, gcap, 128, c->P, S.G, st, c->pos[a].p, c->vel[c->vcur].p, c->kappa[a].p, c->kappav[a].p, c->pid[a].p,
         c->pstate[a].p, X.cap, X.peer_lo ? xchg_rows(X.peer_lo, 1, 0, X.cap) : nul, X.peer_lo ? xchg_rows(X.peer_lo, 1, 1, X.cap) : nul,
         X.peer_lo ? xchg_rows(X.peer_lo, 1, 2, X.cap) : nul, X.peer_hi ? xchg_rows(X.peer_hi, 0, 0, X.cap) : nul,
         X.peer_hi ? xchg_rows(X.peer_hi, 0, 1, X.cap) : nul, X.peer_hi ? xchg_rows(X.peer_hi, 0, 2, X.cap) : nul, S.counts.p, &st->error_flags);
  LAUNCH(c, k_slab_xchg_post_rows, 1, 32, S.counts.p, X, flo, fhi, (volatile unsigned long long *)S.flags.p, S.seq.p + 0, 5000000000ull,
         &st->error_flags);
  LAUNCH(c, k_slab_xchg_unpack, std::max(1, cdiv(2 * (int64_t)X.cap, 128)), 128, st, X, (int)c->nf_cap, c->pos[a].p, c->vel[c->vcur].p,
         c->kappa[a].p, c->kappav[a].p, c->pid[a].p, c->pstate[a].p);
  // sort [own_begin, own_end + received) by (cell, id) into the other buffers
  CU(cudaMemsetAsync(c->cell_start_f.p, 0, sizeof(unsigned int) * (nc + 1), c->ls));
  LAUNCH(c, k_slab_bin_count, gcap, 128, c->P, st, c->pos[a].p, c->cell_start_f.p, c->cell_of_p.p, c->rank_in_cell.p);
  int rc = scan_u32(c, c->cell_start_f.p, (size_t)nc + 1, nullptr);
  if (rc) return rc;
  LAUNCH(c, k_bin_scatter, gcap, 128, (const int *)&st->nf, 0, c->cell_start_f.p, c->cell_of_p.p, c->rank_in_cell.p, c->sorted_src_f.p);
  LAUNCH(c, k_bin_sort_cells_by_id, gcap, 128, (const int *)&st->nf, 0, (const int *)&st->own_begin, c->cell_start_f.p, c->cell_of_p.p,
         c->rank_in_cell.p, c->sorted_src_f.p, c->pid[a].p);
  LAUNCH(c, k_permute_fluid, gcap, 128, st, c->sorted_src_f.p, c->pos[a].p, c->vel[c->vcur].p, c->kappa[a].p, c->kappav[a].p, c->pid[a].p,
         c->pstate[a].p, c->pos[b].p, c->vel[1 - c->vcur].p, c->kappa[b].p, c->kappav[b].p, c->pid[b].p, c->pstate[b].p,
         (const int *)&st->own_begin);
  c->cur = b;
  c->vcur = 1 - c->vcur;
  LAUNCH(c, k_slab_ranges, 1, 32, c->P, S.G, st, c->cell_start_f.p);
  LAUNCH(c, k_slab_xchg_post_layers, 1, 32, st, X, flo, fhi, (volatile unsigned long long *)S.flags.p, S.seq.p + 0, 5000000000ull);
  return DFR_OK;
}

template <bool PRESSURE>
void launch_boundary_side(dfr_context *c, bool grad, int iter_kernel) {
  if (c->n_acc_blocks == 0) return;
  const int g = c->n_acc_blocks, t = BS_WARPS * 32;
  const int a = c->cur;
#define BS_ARGS                                                                                                               \
  c->P, c->dSt.p, c->dBodies.p, c->blk_body.p, c->blk_first.p, c->xk.p, c->vel[c->vcur].p, c->bpos.p, c->bvel.p,         \
      c->bx0.p, c->dyn_begin, c->off_d.p, c->idx_d.p, c->dadv.p, c->factor.p, c->sgp.p, c->pstate[a].p, iter_kernel, \
      c->acc_rows.p
  if (grad) {
    if (PRESSURE)
      LAUNCH(c, (k_boundary_side<0, true>), g, t, BS_ARGS);
    else
      LAUNCH(c, (k_boundary_side<1, true>), g, t, BS_ARGS);
  } else {
    if (PRESSURE)
      LAUNCH(c, (k_boundary_side<0, false>), g, t, BS_ARGS);
    else
      LAUNCH(c, (k_boundary_side<1, false>), g, t, BS_ARGS);
  }
#undef BS_ARGS
}

// divergenceSolve / pressureSolve (TimeStepDiffDFSPH.cpp:770-881 / 654-768).  Iterations are enqueued
// speculatively; every iteration kernel exits at once when the on-device residual test has closed the
// solve, and the host looks at the flag only after the speculated batch.
// fuse_density / fuse_normals (divergence solve with warm start only): the first two k_rho launches of the step also do
// the work of k_density_factor and k_normals (dfr_kernels.cuh: RhoExtra)
// fuse_nonpressure: the last launch of every speculated batch of iterations (the first batch holds as many iterations as
// the previous step needed, every further batch one) also evaluates the non-pressure accelerations;
// *nonpressure_done tells whether the last of those launches really was the last active iteration.
template <bool PRESSURE>
int launch_solver(dfr_context *c, bool fuse_density = false, bool fuse_normals = false, bool fuse_nonpressure = false,
                  bool *nonpressure_done = nullptr) {
  const int n = c->launch_nf, g = cdiv(n, 128);
  const int a = c->cur;
  const bool warm = PRESSURE ? c->cfg.use_pressure_warmstart : c->cfg.use_divergence_warmstart;
  double *kap = PRESSURE ? c->kappa[a].p : c->kappav[a].p;
  RhoExtra X0;
  std::memset(&X0, 0, sizeof(X0));
#define RHO_ARGS_X(POS, EXTRA)                                                                                          \
  c->P, c->dSt.p, (POS), c->vel[c->vcur].p, c->bpos.p, c->bvel.p, list_f(c), list_b(c), c->density.p, c->factor.p, \
      c->pstate[a].p, kap, c->dadv.p, c->xk.p, c->partials.p, ghost_out(c, GA_XK), (EXTRA)
#define RHO_ARGS RHO_ARGS_X(c->pos[a].p, X0)
#define PUSH_ARGS \
  c->P, c->dSt.p, c->xk.p, c->vel[c->vcur].p, c->bpos.p, list_f(c), list_b(c), c->pstate[a].p, kap, warm ? 1 : 0, ghost_out(c, GA_VEL0 + c->vcur)
  if (warm) {
    if (!PRESSURE && fuse_density) {
      RhoExtra X = X0;
      X.density = c->density.p;
      X.factor = c->factor.p;
      X.sgp = c->sgp.p;
      X.xrho = c->xrho.p;
      X.go = ghost_out(c, GA_XRHO);
      PLAUNCH(c, (k_rho<false, RHO_WARM, RHO_X_DENSITY>), g, RHO_ARGS_X(c->pos[a].p, X));
      // peer stores: the flag of the xk update below also covers the (x, rho) rows
      if (!c->slab.p2p) SLAB_SYNC(c, c->xrho.p, sizeof(double4));
    } else
      PLAUNCH(c, (k_rho<PRESSURE, RHO_WARM>), g, RHO_ARGS);
    SLAB_SYNC(c, c->xk.p, sizeof(double4));
    launch_boundary_side<PRESSURE>(c, false, 0);
    PLAUNCH(c, (k_push<PRESSURE, false>), g, PUSH_ARGS);
    SLAB_SYNC(c, c->vel[c->vcur].p, sizeof(double4));
  }
  if (!PRESSURE && warm && fuse_normals) {
    RhoExtra X = X0;
    X.normal = c->normal.p;
    X.go = ghost_out(c, GA_NORMAL);
    PLAUNCH(c, (k_rho<false, RHO_PLAIN, RHO_X_NORMALS>), g, RHO_ARGS_X(c->xrho.p, X));  // (x, rho) records stand in for x
    if (!c->slab.p2p) SLAB_SYNC(c, c->normal.p, sizeof(double4));
  } else
    PLAUNCH(c, (k_rho<PRESSURE, RHO_PLAIN>), g, RHO_ARGS);
  SLAB_SYNC(c, c->xk.p, sizeof(double4));
  const int max_it = PRESSURE ? c->cfg.max_iterations : c->cfg.max_iterations_v;
  if (c->capturing) {
    // graph stepping: the Jacobi iteration is the body of a WHILE node whose condition k_residual_finish writes - no
    // speculation, no read-back.  With the fused non-pressure pass enabled both k_rho variants are in the body and the
    // device decides per iteration which one runs (StepState::fuse_now).
    unsigned long long cond = 0;
    int64_t *outer_counter = c->cap_counter;
    int rc = cap_make_handle(c, 1u, &cond);  // both solves run at least one iteration
    if (!rc) rc = cap_begin_conditional(c, cudaGraphCondTypeWhile, cond);
    if (rc) return rc;
    c->cap_counter = PRESSURE ? &c->cap_target->n_prs_body : &c->cap_target->n_div_body;
    int64_t *outer_syncs = c->slab.cap_sync_counter;
    c->slab.cap_sync_counter = PRESSURE ? &c->slab.cap_syncs_prs : &c->slab.cap_syncs_div;
    launch_boundary_side<PRESSURE>(c, true, 1);
    PLAUNCH(c, (k_push<PRESSURE, true>), g, PUSH_ARGS);
    SLAB_SYNC(c, c->vel[c->vcur].p, sizeof(double4));
    const bool fuse_np = !PRESSURE && fuse_nonpressure;
    if (fuse_np) {
      RhoExtra X = X0;
      X.normal = c->normal.p;
      X.acc = c->acc.p;
      X.gate = 1;
      PLAUNCH(c, (k_rho<false, RHO_ITER, RHO_X_NONPRESSURE>), g, RHO_ARGS_X(c->xrho.p, X));
      RhoExtra Xp = X0;
      Xp.gate = 2;
      PLAUNCH(c, (k_rho<PRESSURE, RHO_ITER>), g, RHO_ARGS_X(c->pos[a].p, Xp));
    } else
      PLAUNCH(c, (k_rho<PRESSURE, RHO_ITER>), g, RHO_ARGS);
    if (c->slab.on) {  // the stopping rule needs the residual of all slabs: summed over peer memory by the deciding kernel
      LAUNCH(c, k_residual_finish<PRESSURE>, RES_BLOCKS, RES_THREADS, c->P, c->dSt.p, c->partials.p, c->partials.p + c->partials.n - RES_BLOCKS,
             0ull, 0);
      LAUNCH(c, k_slab_residual_decide<PRESSURE>, 1, 64, c->P, c->dSt.p, c->slab.mail, cond, fuse_np ? 1 : 0);
    } else
      LAUNCH(c, k_residual_finish<PRESSURE>, RES_BLOCKS, RES_THREADS, c->P, c->dSt.p, c->partials.p, c->partials.p + c->partials.n - RES_BLOCKS,
             cond, fuse_np ? 1 : 0);
    c->cap_counter = outer_counter;
    c->slab.cap_sync_counter = outer_syncs;
    // the two gated k_rho launches count as one executed kernel per iteration
    if (fuse_np) c->cap_target->n_div_body -= 1;
    rc = cap_end_conditional(c);
    (void)max_it;
    return rc;
  }
  int launched = 0;
  int spec = PRESSURE ? c->spec_prs : c->spec_div;
  int fused_at = -1;  // iteration count at which the fused non-pressure pass ran
  const int first_batch = std::max(1, std::min(spec, max_it));
  for (;;) {
    spec = std::max(1, std::min(spec, max_it - launched));
    for (int it = 0; it < spec; it++) {
      c->prof_solver = PRESSURE ? 1 : 0;
      c->prof_iter = launched + it;
      launch_boundary_side<PRESSURE>(c, true, 1);
      PLAUNCH(c, (k_push<PRESSURE, true>), g, PUSH_ARGS);
      SLAB_SYNC(c, c->vel[c->vcur].p, sizeof(double4));
      // the last launch of a batch may be the last iteration: always in a follow-up batch (the solve did not converge in
      // the predicted count, every further iteration probably is the last), in the first batch only while the
      // prediction has been right lately - when the count alternates (1, 2, 1, 2, ...) a fused pass in the first batch is
      // wasted or idle every time
      if (!PRESSURE && fuse_nonpressure && it == spec - 1 && (launched > 0 || c->div_pred_streak >= 2)) {
        RhoExtra X = X0;
        X.normal = c->normal.p;
        X.acc = c->acc.p;
        PLAUNCH(c, (k_rho<false, RHO_ITER, RHO_X_NONPRESSURE>), g, RHO_ARGS_X(c->xrho.p, X));
        fused_at = launched + spec;
      } else
        PLAUNCH(c, (k_rho<PRESSURE, RHO_ITER>), g, RHO_ARGS);
      LAUNCH(c, k_residual_finish<PRESSURE>, RES_BLOCKS, RES_THREADS, c->P, c->dSt.p, c->partials.p, c->partials.p + c->partials.n - RES_BLOCKS,
             0ull, 0);
      if (c->slab.on && c->slab.mail_ok) {  // the residual of the iteration is the sum over all slabs: one kernel over peer memory
        LAUNCH(c, k_slab_residual_decide<PRESSURE>, 1, 64, c->P, c->dSt.p, c->slab.mail, 0ull, 0);
      } else if (c->slab.on) {
        int rc = slab_allreduce(c, &c->dSt.p->res_sum, 1, ncclDouble, ncclSum);
        if (rc) return rc;
        LAUNCH(c, k_solver_decide<PRESSURE>, 1, 32, c->P, c->dSt.p);
        // peer-store transport: every rank's k_rho (and with it its stores into my ghost range) precedes its
        // contribution to the all-reduce in stream order, so the completed all-reduce already orders them before my
        // next kernel; with NCCL send/recv the rows still have to be shipped
        if (!c->slab.p2p) SLAB_SYNC(c, c->xk.p, sizeof(double4));
      }
    }
    c->prof_solver = -1;
    launched += spec;
    int rc = sync_state(c);
    if (rc) return rc;
    if (c->profiling) prof_resolve(c, c->hSt->div_iters, c->hSt->prs_iters);
    const int active = PRESSURE ? c->hSt->prs_active : c->hSt->div_active;
    if (!active || launched >= max_it) break;
    // the prediction was too low: follow-up batches of 1, 2, 4, 8, 8, ... iterations (a read-back per batch; launches
    // past convergence return at once)
    spec = (launched <= first_batch) ? 1 : std::min(2 * spec, 8);
  }
  const int used = PRESSURE ? c->hSt->prs_iters : c->hSt->div_iters;
  if (nonpressure_done) *nonpressure_done = (fused_at > 0 && used == fused_at && getenv_int("DFR_NO_FUSION") != 3);
  if (PRESSURE)
    c->spec_prs = std::max(used, c->cfg.min_iterations);
  else {
    c->div_pred_streak = (used == c->spec_div) ? std::min(c->div_pred_streak + 1, 1000) : 0;
    c->spec_div = std::max(used, 1);
  }
#undef RHO_ARGS
#undef PUSH_ARGS
  return DFR_OK;
}

// ---- contact order: the reference applies the contact impulses in its storage order, which CompactNSearch's z_sort
// (a stable sort by the Morton code of floor(x / support radius)) permutes at deferredInit (body-frame positions), at
// every reset (positions left by the previous trajectory) and every 500 steps; see oracle/oracle_contact.inc ----
void contact_sort_with(dfr_context *c, const std::vector<double4> &pos) {
  const double r = c->P.support_radius;
  auto spread3 = [](uint64_t v) {
    uint64_t o = 0;
    for (int k = 0; k < 21; k++) o |= ((v >> k) & 1ull) << (3 * k);
    return o;
  };
  const long long B = 1ll << 20;
  std::vector<uint64_t> code(pos.size());
  for (size_t i = 0; i < pos.size(); i++) {
    const long long cx = (long long)std::floor(pos[i].x / r), cy = (long long)std::floor(pos[i].y / r), cz = (long long)std::floor(pos[i].z / r);
    code[i] = spread3((uint64_t)(cx + B)) | (spread3((uint64_t)(cy + B)) << 1) | (spread3((uint64_t)(cz + B)) << 2);
  }
  for (auto &hb : c->bodies) {
    if (!hb.dynamic) continue;
    const int first = hb.dev0.p_begin - c->dyn_begin;
    std::stable_sort(c->h_order.begin() + first, c->h_order.begin() + first + hb.n, [&](int a, int b) { return code[a] < code[b]; });
  }
}
int contact_upload_order(dfr_context *c) {
  if (c->n_dyn_p == 0) return DFR_OK;
  CU(cudaMemcpyAsync(c->c_order.p, c->h_order.data(), c->n_dyn_p * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return DFR_OK;
}
int contact_sort_current(dfr_context *c) {  // keys from the current world positions of the dynamic particles
  if (c->n_dyn_p == 0) return DFR_OK;
  c->h_dynpos.resize(c->n_dyn_p);
  CU(cudaMemcpyAsync(c->h_dynpos.data(), c->bpos.p + c->dyn_begin, c->n_dyn_p * sizeof(double4), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  contact_sort_with(c, c->h_dynpos);
  return contact_upload_order(c);
}
// RigidContactSolver ctor (:23-263): rest volumes / densities, then the first z-sort (body-frame positions)
int contact_init(dfr_context *c) {
  const size_t NB = (size_t)std::max(c->n_b, 1), ND = (size_t)std::max(c->n_dyn_p, 1);
  CU(c->c_vol0.alloc(NB)); CU(c->c_dens0.alloc(NB)); CU(c->c_dens.alloc(ND)); CU(c->c_vel.alloc(ND)); CU(c->c_order.alloc(ND));
  CU(c->c_records.alloc(ND * CREC_N));
  const double hc = c->cfg.rigid_contact_support_radius_factor * c->cfg.particle_radius;
  c->CP.inv_h = 1.0 / hc;
  c->CP.k_cubic = 8.0 / (M_PI * hc * hc * hc);
  c->CP.gamma = c->cfg.rigid_contact_gamma;
  c->CP.beta = c->cfg.rigid_contact_beta;
  c->CP.mu = c->cfg.rigid_contact_friction;
  if (hc > c->P.support_radius * (1.0 + 1e-12))
    return fail(c, DFR_ERR_INVALID, "rigidContactSupportRadiusFactor above 4: the reference searches contacts with the fluid support radius");
  c->h_order.resize(c->n_dyn_p);
  std::vector<double4> local(c->n_dyn_p);
  for (auto &hb : c->bodies) {
    if (!hb.dynamic) continue;
    const int first = hb.dev0.p_begin - c->dyn_begin;
    for (int64_t j = 0; j < hb.n; j++) {
      c->h_order[first + j] = first + (int)j;
      local[first + j] = make_double4(hb.x_local[3 * j], hb.x_local[3 * j + 1], hb.x_local[3 * j + 2], 0.0);
    }
  }
  contact_sort_with(c, local);
  return DFR_OK;
}
// vol0 / density0 from the (rigidly moved, hence equal up to rounding) world-space samples; needs the cell tables
int contact_rest_state(dfr_context *c) {
  if (c->n_b == 0) return DFR_OK;
  for (int pass = 0; pass < 2; pass++)
    LAUNCH(c, k_contact_rest, cdiv(c->n_b, 128), 128, c->P, c->CP, pass, c->bpos.p, c->bbody.p, c->n_b, c->n_static_p, grid_static(c),
           grid_dyn(c), c->n_static_p > 0 ? 1 : 0, c->n_dyn_p > 0 ? 1 : 0, c->c_vol0.p, c->c_dens0.p);
  return DFR_OK;
}

// one SimulatorBase::timeStepNoGUI body (SimulatorBase.cpp:1142-1169)
// Neighbour rows that do not fit their ELL capacity: the reference has no such limit (vector<vector<unsigned>>), so the
// capacities grow and the lists are rebuilt - the sort is done and nothing else of the step has run yet.  The check
// needs a host read-back right after the list build, so it is only made while rows may plausibly overflow: in the
// first steps after finalize / reset / load, and whenever the longest row of the previous step was above half the
// capacity (a row does not double within one step of a CFL-limited simulation).  Otherwise an overflow still surfaces
// as DFR_ERR_CAPACITY at the step's first read-back.  Slab-decomposed contexts only come here in the steps after
// finalize / reset / load (enqueue_steps counts them down) and under profiling: always checked; the lists are local, so
// a rank that rebuilds its lists changes nothing its neighbours can see.
int ensure_list_capacity(dfr_context *c) {
  const bool risky = c->slab.on || c->fresh_steps > 0 || 2 * (int)c->hSt->list_used_f > c->cap_f ||
                     2 * (int)c->hSt->list_used_b > c->cap_b || 2ull * c->hSt->list_used_d > c->cap_d;
  if (!c->slab.on && c->fresh_steps > 0) c->fresh_steps--;
  if (!risky) return DFR_OK;
  for (int attempt = 0; attempt < 6; attempt++) {
    CU(cudaMemcpyAsync(c->hSt, c->dSt.p, sizeof(StepState), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (c->hSt->error_flags & ~7) break;  // a slab error: reported below, nothing to grow
    const int flags = c->hSt->error_flags & 7;
    if (!flags) return DFR_OK;
    const size_t nwarp = ((size_t)c->nf_cap + 31) / 32;
    if (flags & 1) {
      const int want = (std::max(2 * c->cap_f, (int)(c->hSt->list_used_f * 5 / 4) + 8) + 3) & ~3;
      if (want > 4096) break;
      c->idx_f.free();
      CU(c->idx_f.alloc(nwarp * 32 * (size_t)want));
      c->cap_f = want;
    }
    if (flags & 2) {
      const int want = (std::max(2 * c->cap_b, (int)(c->hSt->list_used_b * 5 / 4) + 8) + 3) & ~3;
      if (want > 4096) break;
      c->idx_b.free();
      CU(c->idx_b.alloc(nwarp * 32 * (size_t)want));
      c->cap_b = want;
    }
    if (flags & 4) {
      const unsigned long long want = std::max<unsigned long long>(2ull * c->cap_d, (unsigned long long)c->hSt->list_used_d * 5 / 4 + 1024);
      if (want > (1ull << 31)) break;
      c->idx_d.free();
      CU(c->idx_d.alloc((size_t)want));
      c->cap_d = (unsigned int)want;
    }
    CU(cudaMemsetAsync(&c->dSt.p->error_flags, 0, sizeof(int), c->stream));
    int rc = build_neighbor_lists(c);
    if (rc) return rc;
  }
  return sync_state(c);  // reports the capacity error
}

// TimeStepDiffDFSPH::performNeighborhoodSearch (:2044-2056): z-sort of the contact order every 500 steps (host side:
// a read-back of the dynamic particles, so it stays outside the step graph)
int contact_sort_tick(dfr_context *c) {
  if (!c->cfg.use_rigid_contact_solver) return DFR_OK;
  if (c->sort_counter % 500 == 0) {
    int rc = contact_sort_current(c);
    if (rc) return rc;
  }
  c->sort_counter++;
  return DFR_OK;
}
bool fuse_nonpressure_enabled(const dfr_context *c) {
  const int no_fusion = getenv_int("DFR_NO_FUSION");
  return c->cfg.enable_divergence_solver && c->cfg.use_divergence_warmstart && no_fusion != 1 && no_fusion != 2;
}

// Enqueues one step on c->ls.  While a step graph is being recorded (c->capturing) the host-side parts - contact order
// sort, capacity watch - are left to the caller and the solver loops become WHILE nodes.
int launch_step(dfr_context *c) {
  int n = c->launch_nf, g = cdiv(n, 128);
  int rc = DFR_OK;
  if (!c->capturing) {
    rc = contact_sort_tick(c);
    if (rc) return rc;
  }
  if (c->capturing && c->slab.on && !c->slab.xchg_ok)
    // DFR_SLAB_HOST_EXCHANGE=1: the step is replayed from the list build on; k_begin_step and the particle exchange (NCCL
    // transfers whose sizes the host reads back) stay on the stream, see enqueue_steps
    rc = build_neighbor_lists(c);
  else {
    LAUNCH(c, k_begin_step, 1, 32, c->P, c->dSt.p, c->dBodies.p, (c->capturing && fuse_nonpressure_enabled(c)) ? 1 : 0);
    rc = build_neighbors(c);
  }
  if (rc) return rc;
  if (!c->capturing) {
    rc = ensure_list_capacity(c);
    if (rc) return rc;
  }
  n = c->launch_nf;  // slab mode: the number of local particles changes with every exchange
  g = cdiv(n, 128);
  int a = c->cur;
  // with the divergence solve and its warm start on, density/factor, the normals and the non-pressure accelerations ride
  // on k_rho passes of that solve (dfr_kernels.cuh: RhoExtra); DFR_NO_FUSION=1 switches all of it off, =2 only the last,
  // =3 (tests) runs the fused non-pressure pass but always discards it in favour of k_nonpressure
  const int no_fusion = getenv_int("DFR_NO_FUSION");
  const bool fuse = c->cfg.enable_divergence_solver && c->cfg.use_divergence_warmstart && no_fusion != 1;
  const bool fuse_normals = fuse && c->cfg.surface_tension_method == 2;
  const bool fuse_nonpressure = fuse && no_fusion != 2;
  if (!fuse) {
    PLAUNCH(c, k_density_factor, g, c->P, c->dSt.p, c->pos[a].p, c->bpos.p, list_f(c), list_b(c), c->density.p, c->factor.p,
           c->sgp.p, c->xrho.p, ghost_out(c, GA_XRHO));
    // x|rho is first gathered by k_normals / k_nonpressure: with peer stores any later flag (the divergence solve has
    // several) covers it
    if (!(c->slab.p2p && c->cfg.enable_divergence_solver)) SLAB_SYNC(c, c->xrho.p, sizeof(double4));
  }
  bool scale_kv = false, nonpressure_done = false;
  if (c->cfg.enable_divergence_solver) {
    rc = launch_solver<false>(c, fuse, fuse_normals, fuse_nonpressure, &nonpressure_done);
    if (rc) return rc;
    scale_kv = c->cfg.use_divergence_warmstart != 0;
  }
  // reaction of the boundary viscosity on dynamic bodies (force, torque, dF/dv), from the velocities the non-pressure pass sees
  if (c->cfg.viscosity_method == 1 && c->cfg.viscosity_boundary != 0.0 && c->n_acc_blocks > 0)
    LAUNCH(c, k_boundary_viscosity, c->n_acc_blocks, BS_WARPS * 32, c->P, c->dSt.p, c->dBodies.p, c->blk_body.p, c->blk_first.p, c->xrho.p,
           c->vel[c->vcur].p, c->bpos.p, c->bvel.p, c->dyn_begin, c->off_d.p, c->idx_d.p, c->acc_rows.p);
  if (c->cfg.surface_tension_method == 2 && !fuse_normals)
  {
    PLAUNCH(c, k_normals, g, c->P, c->dSt.p, c->xrho.p, list_f(c), c->normal.p, ghost_out(c, GA_NORMAL));
    SLAB_SYNC(c, c->normal.p, sizeof(double4));
  }
  if (c->capturing && fuse_nonpressure && c->cfg.enable_divergence_solver) {
    // graph stepping: both are enqueued, StepState::np_done picks the one that runs (counted as one launch)
    PLAUNCH(c, k_nonpressure, g, c->P, c->dSt.p, c->xrho.p, c->vel[c->vcur].p, c->bpos.p, c->bvel.p, list_f(c), list_b(c),
           c->normal.p, c->pstate[a].p, c->kappav[a].p, scale_kv ? 1 : 0, c->acc.p, c->vel[1 - c->vcur].p,
           ghost_out(c, GA_VEL0 + (1 - c->vcur)), 1);
    LAUNCH(c, k_apply_accel, g, 128, c->dSt.p, c->acc.p, c->vel[c->vcur].p, c->pstate[a].p, c->kappav[a].p, scale_kv ? 1 : 0,
           c->vel[1 - c->vcur].p, ghost_out(c, GA_VEL0 + (1 - c->vcur)), 2);
    (*c->cap_counter)--;
  } else if (nonpressure_done)
    LAUNCH(c, k_apply_accel, g, 128, c->dSt.p, c->acc.p, c->vel[c->vcur].p, c->pstate[a].p, c->kappav[a].p, scale_kv ? 1 : 0,
           c->vel[1 - c->vcur].p, ghost_out(c, GA_VEL0 + (1 - c->vcur)), 0);
  else
    PLAUNCH(c, k_nonpressure, g, c->P, c->dSt.p, c->xrho.p, c->vel[c->vcur].p, c->bpos.p, c->bvel.p, list_f(c), list_b(c),
           c->normal.p, c->pstate[a].p, c->kappav[a].p, scale_kv ? 1 : 0, c->acc.p, c->vel[1 - c->vcur].p,
           ghost_out(c, GA_VEL0 + (1 - c->vcur)), 0);
  c->vcur = 1 - c->vcur;
  if (!c->slab.p2p) SLAB_SYNC(c, c->vel[c->vcur].p, sizeof(double4));  // peer stores: ordered by the CFL all-reduce below
  if (c->n_dyn_p > 0) LAUNCH(c, k_cfl_boundary, cdiv(c->n_dyn_p, 128), 128, c->dSt.p, c->bvel.p, c->dyn_begin, c->n_dyn_p);
  if (c->slab.on && c->slab.mail_ok)  // max |v + a h|^2 over all slabs (ordered bits of positive doubles)
    LAUNCH(c, k_slab_max_u64, 1, 64, c->dSt.p, c->slab.mail, &c->dSt.p->cfl_max_bits);
  else if (c->slab.on) {
    rc = slab_allreduce(c, &c->dSt.p->cfl_max_bits, 1, ncclUint64, ncclMax);
    if (rc) return rc;
  }
  LAUNCH(c, k_cfl_finish, 1, 32, c->P, c->dSt.p);
  rc = launch_solver<true>(c);
  if (rc) return rc;
  LAUNCH(c, k_advect_x, g, 128, c->dSt.p, c->pos[a].p, c->vel[c->vcur].p, c->pstate[a].p, c->kappa[a].p,
         c->cfg.use_pressure_warmstart ? 1 : 0);
  if (!c->h_emitters.empty()) {  // emitParticles (TimeStepDiffDFSPH.cpp:637-641)
    LAUNCH(c, k_emit_release, g, 128, c->dSt.p, c->pstate[a].p);
    for (int ei = 0; ei < (int)c->h_emitters.size(); ei++) {
      LAUNCH(c, k_emit_animate, g, 128, c->P, c->dSt.p, c->dEmitters.p, ei, c->pos[a].p, c->vel[c->vcur].p, c->pstate[a].p);
      LAUNCH(c, k_emit_spawn, 1, 256, c->P, c->dSt.p, c->dEmitters.p, ei, (int)c->nf_cap, c->pos[a].p, c->vel[c->vcur].p, c->kappa[a].p,
             c->kappav[a].p, c->pid[a].p, c->pstate[a].p);
    }
  }
  if (c->P.n_bodies > 0) {
    if (c->slab.on) {  // per-body force / torque / Jacobian rows: sum over the slabs, then every rank advances the bodies alike
      LAUNCH(c, k_body_rows_to_buf, c->P.n_bodies, 256, c->dBodies.p, c->acc_rows.p, c->slab.body_buf.p);
      if (c->slab.mail_ok)
        LAUNCH(c, k_slab_sum_f64, 1, 256, c->dSt.p, c->slab.mail, c->slab.body_buf.p, c->P.n_bodies * ACC_N);
      else {
        rc = slab_allreduce(c, c->slab.body_buf.p, (size_t)c->P.n_bodies * ACC_N, ncclDouble, ncclSum);
        if (rc) return rc;
      }
      LAUNCH(c, k_body_buf_apply, c->P.n_bodies, 32, c->dBodies.p, c->slab.body_buf.p);
    } else
      LAUNCH(c, k_body_reduce, c->P.n_bodies, 256, c->dBodies.p, c->acc_rows.p);
    if (c->cfg.use_rigid_contact_solver) {
      LAUNCH(c, k_body_update, 1, 32, c->P, c->dSt.p, c->dBodies.p, c->dMgr.p, (int)BODY_PRE);
      if (c->n_dyn_p > 0) {
        const int gd = cdiv(c->n_dyn_p, 128);
        LAUNCH(c, k_contact_prepare, gd, 128, c->P, c->CP, c->dBodies.p, c->bpos.p, c->bbody.p, c->dyn_begin, c->n_dyn_p, c->n_static_p,
               grid_static(c), grid_dyn(c), c->n_static_p > 0 ? 1 : 0, c->c_vol0.p, c->c_vel.p, c->c_dens.p);
        LAUNCH(c, k_contact_force, gd, 128, c->P, c->CP, c->dBodies.p, c->bpos.p, c->bx0.p, c->bbody.p, c->dyn_begin, c->n_dyn_p,
               c->n_static_p, grid_static(c), grid_dyn(c), c->n_static_p > 0 ? 1 : 0, c->c_vol0.p, c->c_dens0.p, c->c_vel.p, c->c_dens.p,
               c->c_records.p);
        if (c->profiling) prof_begin(c, "k_contact_apply");
        k_contact_apply<<<c->P.n_bodies, 32, (1 + c->P.n_bodies) * CG_N * sizeof(double), c->ls>>>(
            c->P, c->dSt.p, c->dBodies.p, c->dMgr.p, c->bpos.p, c->dyn_begin, c->c_order.p, c->c_records.p);
        if (c->profiling) prof_end(c);
        if (c->capturing) (*c->cap_counter)++;
        else c->launches++;
      }
      LAUNCH(c, k_body_update, 1, 32, c->P, c->dSt.p, c->dBodies.p, c->dMgr.p, (int)BODY_POST);
    } else {
      LAUNCH(c, k_body_update, 1, 32, c->P, c->dSt.p, c->dBodies.p, c->dMgr.p, (int)BODY_ALL);
    }
    if (c->n_dyn_p > 0)
      LAUNCH(c, k_update_boundary_particles, cdiv(c->n_dyn_p, 128), 128, c->dBodies.p, c->bbody.p, c->bx0.p, c->bpos.p, c->bvel.p,
             c->dyn_begin, c->n_dyn_p, 0);
  } else {
    LAUNCH(c, k_body_update, 1, 32, c->P, c->dSt.p, c->dBodies.p, c->dMgr.p, (int)BODY_ALL);
  }
  return DFR_OK;
}

int reset_device_state(dfr_context *c) {
  c->bodies_mirrored = false;
  if (c->cfg.use_rigid_contact_solver) {
    if (!c->contact_ready) {
      int rc = contact_init(c);
      if (rc) return rc;
    } else {  // RigidBody3dBoundarySimulator::reset (:346-362): z-sort before the particles return to their start pose
      int rc = contact_sort_current(c);
      if (rc) return rc;
    }
    c->sort_counter = 0;
  }
  // fluid: initial state back into the current buffers (id order)
  const size_t n = (size_t)c->nf_loc0;
  c->cur = 0;
  c->vcur = 0;
  if (n) {
    CU(cudaMemcpyAsync(c->pos[0].p, c->pos_init.p, n * sizeof(double4), cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(c->vel[0].p, c->vel_init.p, n * sizeof(double4), cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(c->kappa[0].p, c->kappa_init.p, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(c->kappav[0].p, c->kappav_init.p, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    LAUNCH(c, k_iota_i32, cdiv(c->nf_cap, 256), 256, c->pid[0].p, (size_t)c->nf_cap);  // particle ids = input order
    if (c->slab.on) {  // a slab holds a subset: its ids at t = 0
      CU(cudaMemcpyAsync(c->pid[0].p, c->h_ids0.data(), c->h_ids0.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
      CU(cudaStreamSynchronize(c->stream));
    }
    CU(cudaMemsetAsync(c->pstate[0].p, 0, c->nf_cap * sizeof(int), c->stream));
    CU(cudaMemsetAsync(c->sgp.p, 0, c->nf_cap * sizeof(double4), c->stream));
    CU(cudaMemsetAsync(c->acc.p, 0, c->nf_cap * sizeof(double4), c->stream));
    CU(cudaMemsetAsync(c->density.p, 0, c->nf_cap * sizeof(double), c->stream));
  }
  // bodies
  if (!c->bodies.empty()) {
    std::vector<BodyDev> hb(c->bodies.size());
    for (size_t i = 0; i < c->bodies.size(); i++) {
      host_body_record(c, c->bodies[i], c->bodies[i].dev0.p_begin);
      hb[i] = c->bodies[i].dev0;
    }
    // keep block slices
    CU(cudaMemcpyAsync(c->dBodies.p, hb.data(), hb.size() * sizeof(BodyDev), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    std::vector<MgrBlock> mg(c->bodies.size() * c->bodies.size());
    std::memset(mg.data(), 0, mg.size() * sizeof(MgrBlock));
    for (size_t R = 0; R < c->bodies.size(); R++) {  // RigidBodyGradientManager::reset (:477-524)
      mg[R * c->bodies.size() + R].vn_v0 = m33::identity();
      mg[R * c->bodies.size() + R].wn_w0 = m33::identity();
    }
    CU(cudaMemcpyAsync(c->dMgr.p, mg.data(), mg.size() * sizeof(MgrBlock), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (c->acc_rows.n) CU(cudaMemsetAsync(c->acc_rows.p, 0, c->acc_rows.n * sizeof(double), c->stream));
  }
  if (!c->h_emitters.empty())  // Emitter::reset: next emit time back to the start
    CU(cudaMemcpyAsync(c->dEmitters.p, c->h_emitters.data(), c->h_emitters.size() * sizeof(EmitterDev), cudaMemcpyHostToDevice, c->stream));
  StepState st;
  std::memset(&st, 0, sizeof(st));
  st.h = c->cfg.time_step_size;
  st.h_step = st.h;
  st.nf = (int)c->nf_loc0;
  st.own_begin = 0;
  st.own_end = st.nf;
  st.spec_div = 1;
  st.spec_div_prev = 1;
  st.div_streak = 2;
  if (c->slab.on) {
    c->slab.h_ranges[0] = 0;
    c->slab.h_ranges[1] = c->slab.h_ranges[2] = c->slab.h_ranges[3] = c->slab.h_ranges[4] = st.nf;
    c->slab.exchanged_bytes = 0;
    c->slab.h_stale = false;
    c->launch_nf = st.nf;
  }
  CU(cudaMemcpyAsync(c->dSt.p, &st, sizeof(st), cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  *c->hSt = st;
  // boundary particles in world space (dynamic ones; static never move), psi
  if (c->n_dyn_p > 0)
    LAUNCH(c, k_update_boundary_particles, cdiv(c->n_dyn_p, 128), 128, c->dBodies.p, c->bbody.p, c->bx0.p, c->bpos.p, c->bvel.p,
           c->dyn_begin, c->n_dyn_p, 1);
  int rc = build_dyn_grid(c);
  if (rc) return rc;
  rc = compute_boundary_volumes(c);
  if (rc) return rc;
  if (c->cfg.use_rigid_contact_solver) {
    if (!c->contact_ready) {
      rc = contact_rest_state(c);
      if (rc) return rc;
      c->contact_ready = true;
    }
    rc = contact_upload_order(c);
    if (rc) return rc;
  }
  CU(cudaStreamSynchronize(c->stream));
  c->spec_div = 1;
  c->div_pred_streak = 2;
  c->fresh_steps = 4;
  c->spec_prs = std::max(2, c->cfg.min_iterations);
  c->device_ms = 0.0;
  c->launches = 0;
  return DFR_OK;
}

}  // namespace

template <int R, int C>
static void put(double *out, const Mat<R, C> &m) {
  for (int k = 0; k < R * C; k++) out[k] = m.a[k];
}

// ===============================================================================================
// (init_v, init_omega) values staged by dfr_set_init_v_omega -> device, in stream order
int upload_staged_init(dfr_context *c) {
  if (!c->init_dirty) return DFR_OK;
  static_assert(offsetof(BodyDev, init_omega) == offsetof(BodyDev, init_v) + sizeof(d3), "init_v and init_omega are adjacent");
  for (size_t b = 0; b < c->bodies.size(); b++) {
    if (!c->bodies[b].init_staged) continue;
    c->bodies[b].init_staged = false;
    CU(cudaMemcpyAsync(&(c->dBodies.p + b)->init_v, c->h_init + 6 * b, 2 * sizeof(d3), cudaMemcpyHostToDevice, c->stream));
  }
  c->init_dirty = false;
  return DFR_OK;
}

// Shared-memory carve-out hint for the gather kernels that own static shared memory (residual / CFL reductions): they
// live off the L1, but too small a carve-out limits their residency.  DFR_CARVEOUT_PCT (tuning): percent of the
// unified L1/shared array asked for, -1 = leave the driver's default.
template <class K>
void carveout_hint(K kernel, int pct) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}
void set_cache_preferences() {
  const char *v = std::getenv("DFR_CARVEOUT_PCT");
  const int pct = (v && *v) ? std::atoi(v) : -1;
  if (pct < 0) return;
  carveout_hint(k_rho<false, RHO_ITER, RHO_X_NONE>, pct);
  carveout_hint(k_rho<true, RHO_ITER, RHO_X_NONE>, pct);
  carveout_hint(k_rho<false, RHO_ITER, RHO_X_NONPRESSURE>, pct);
  carveout_hint(k_nonpressure, pct);
  cudaGetLastError();  // the hint is best effort
}

extern "C" {

void dfr_default_config(dfr_config *cfg) {
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->particle_radius = 0.025;
  cfg->density0 = 1000.0;
  cfg->gravitation[1] = -9.81;
  cfg->cfl_method = 1;
  cfg->cfl_factor = 0.5;
  cfg->cfl_min_time_step = 0.0001;
  cfg->cfl_max_time_step = 0.005;
  cfg->time_step_size = 0.001;
  cfg->min_iterations = 2;
  cfg->max_iterations = 100;
  cfg->max_error = 0.01;
  cfg->max_iterations_v = 100;
  cfg->max_error_v = 0.1;
  cfg->enable_divergence_solver = 1;
  cfg->use_pressure_warmstart = 1;
  cfg->use_divergence_warmstart = 1;
  cfg->viscosity_method = 1;
  cfg->viscosity = 0.01;
  cfg->surface_tension_method = 0;
  cfg->surface_tension = 0.05;
  cfg->gradient_mode = 1;
  cfg->rigid_body_mode = 0;
  cfg->optimize_rotation = 1;
  cfg->rigid_contact_beta = 1.0;
  cfg->rigid_contact_gamma = 0.7;
  cfg->rigid_contact_support_radius_factor = 4.0;
  cfg->target_time = 0.8;
}

int dfr_create(const dfr_config *cfg, int device, dfr_context **out) {
  if (!cfg || !out) return DFR_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return DFR_ERR_NO_DEVICE;  // no CPU fallback
  if (device < 0 || device >= ndev) return DFR_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return DFR_ERR_CUDA;
  dfr_context *c = new dfr_context();
  c->cfg = *cfg;
  c->device = device;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&c->ev0) != cudaSuccess ||
      cudaEventCreate(&c->ev1) != cudaSuccess || cudaMallocHost((void **)&c->hSt, sizeof(StepState)) != cudaSuccess) {
    delete c;
    return DFR_ERR_CUDA;
  }
  std::memset(c->hSt, 0, sizeof(StepState));
  c->ls = c->stream;
  *out = c;
  return DFR_OK;
}

void dfr_destroy(dfr_context *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  drop_step_graphs(c);
  for (auto &cs : c->cap_stream)
    if (cs) cudaStreamDestroy(cs);
  if (c->side_stream) cudaStreamDestroy(c->side_stream);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  for (int k = 0; k < 2; k++) {
    c->pos[k].free(); c->vel[k].free(); c->kappa[k].free(); c->kappav[k].free(); c->pid[k].free(); c->pstate[k].free();
  }
  c->acc.free(); c->sgp.free(); c->normal.free(); c->density.free(); c->factor.free(); c->dadv.free(); c->xk.free(); c->xrho.free();
  c->partials.free(); c->pos_init.free(); c->vel_init.free(); c->kappa_init.free(); c->kappav_init.free();
  c->bpos.free(); c->bvel.free(); c->bx0.free(); c->bpos_tmp.free(); c->bx0_tmp.free(); c->bbody.free(); c->borig.free();
  c->bbody_tmp.free(); c->borig_tmp.free(); c->bvol.free(); c->dBodies.free(); c->dMgr.free(); c->acc_rows.free();
  c->blk_body.free(); c->blk_first.free(); c->cell_start_f.free(); c->cell_start_s.free(); c->cell_start_d.free();
  c->near_s.free();
  c->near_d.free();
  c->tile_sums.free(); c->tile_sums_side.free(); c->cell_of_p.free(); c->rank_in_cell.free(); c->sorted_src_f.free(); c->sorted_src_d.free();
  c->cell_of_b.free(); c->rank_b.free(); c->cnt_f.free(); c->cnt_b.free(); c->idx_f.free(); c->idx_b.free(); c->idx_d.free();
  c->off_d.free(); c->dSt.free(); c->dEmitters.free();
  c->sched_ctr.free();
  for (int k = 0; k < 2; k++) { c->slab.s_pos[k].free(); c->slab.s_vel[k].free(); c->slab.s_misc[k].free(); }
  for (void *p : c->slab.ipc_opened)
    if (p) cudaIpcCloseMemHandle(p);
  for (void *p : c->slab.ipc_more)
    if (p) cudaIpcCloseMemHandle(p);
  c->slab.seq.free(); c->slab.mbox.free(); c->slab.mflag.free();
  c->slab.r_misc.free(); c->slab.counts.free(); c->slab.xin.free(); c->slab.body_buf.free(); c->slab.flags.free(); c->slab.ipc_stage.free();
  if (c->slab.h_counts) cudaFreeHost(c->slab.h_counts);
  if (c->slab.h_stage) cudaFreeHost(c->slab.h_stage);
  if (c->slab.comm && nccl_api(nullptr)) nccl_api(nullptr)->CommDestroy(c->slab.comm);
  c->c_vol0.free(); c->c_dens0.free(); c->c_dens.free(); c->c_records.free(); c->c_vel.free(); c->c_order.free();
  for (auto &p : c->prof_pending) {
    cudaEventDestroy(p.e0);
    cudaEventDestroy(p.e1);
  }
  for (auto e : c->prof_pool) cudaEventDestroy(e);
  if (c->hSt) cudaFreeHost(c->hSt);
  if (c->h_bodies) cudaFreeHost(c->h_bodies);
  if (c->h_init) cudaFreeHost(c->h_init);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char *dfr_last_error(const dfr_context *c) { return c ? c->err.c_str() : "null context"; }

int dfr_set_fluid(dfr_context *c, int64_t n, const double *x, const double *v) {
  if (!c) return DFR_ERR_INVALID;
  if (c->finalized) return fail(c, DFR_ERR_STATE, "set_fluid after finalize");
  if (n < 0 || (n > 0 && !x)) return fail(c, DFR_ERR_INVALID, "bad fluid arrays");
  c->nf0 = n;
  c->h_fx.assign(x, x + 3 * n);
  if (v)
    c->h_fv.assign(v, v + 3 * n);
  else
    c->h_fv.assign(3 * n, 0.0);
  return DFR_OK;
}

int dfr_add_body(dfr_context *c, int64_t n, const double *x_local, int is_dynamic, double density, const double position[3],
                 const double quat_wxyz[4]) {
  if (!c) return DFR_ERR_INVALID;
  if (c->finalized) return fail(c, DFR_ERR_STATE, "add_body after finalize");
  if (n <= 0 || !x_local || !position || !quat_wxyz) return fail(c, DFR_ERR_INVALID, "bad body arrays");
  if (c->bodies.size() >= 32) return fail(c, DFR_ERR_CAPACITY, "at most 32 rigid bodies per context");
  HostBody hb;
  hb.n = n;
  hb.x_local.assign(x_local, x_local + 3 * n);
  hb.dynamic = is_dynamic ? 1 : 0;
  hb.density = density;
  std::memcpy(hb.pos, position, sizeof(hb.pos));
  std::memcpy(hb.q, quat_wxyz, sizeof(hb.q));
  c->bodies.push_back(std::move(hb));
  return (int)c->bodies.size() - 1;
}

int dfr_set_init_v_omega(dfr_context *c, int body, const double v0[3], const double omega0[3]) {
  if (!c || body < 0 || body >= (int)c->bodies.size()) return fail(c, DFR_ERR_INVALID, "bad body index");
  HostBody &hb = c->bodies[body];
  const bool same = (!v0 || std::memcmp(hb.init_v, v0, sizeof(hb.init_v)) == 0) && (!omega0 || std::memcmp(hb.init_w, omega0, sizeof(hb.init_w)) == 0);
  if (v0) std::memcpy(hb.init_v, v0, sizeof(hb.init_v));
  if (omega0) std::memcpy(hb.init_w, omega0, sizeof(hb.init_w));
  if (c->finalized && !same) {  // SimulationDataDiffDFSPH::get_init_v_rb is read at every beginStep
    hb.dev0.init_v = mk3(hb.init_v[0], hb.init_v[1], hb.init_v[2]);
    hb.dev0.init_omega = mk3(hb.init_w[0], hb.init_w[1], hb.init_w[2]);
    // staged in pinned memory and uploaded in stream order at the start of the next step / trajectory / reset: the
    // control input of a step costs one asynchronous 48-byte copy, no synchronisation
    for (int k = 0; k < 3; k++) {
      c->h_init[6 * body + k] = hb.init_v[k];
      c->h_init[6 * body + 3 + k] = hb.init_w[k];
    }
    hb.init_staged = true;
    c->init_dirty = true;
  }
  return DFR_OK;
}

int dfr_add_emitter(dfr_context *c, int width, int height, const double position[3], const double rot[9], double velocity,
                    double emit_start, double emit_end) {
  if (!c) return DFR_ERR_INVALID;
  if (c->finalized) return fail(c, DFR_ERR_STATE, "add_emitter after finalize");
  if (width <= 0 || height <= 0 || !position || !rot || !(velocity > 0.0)) return fail(c, DFR_ERR_INVALID, "bad emitter");
  EmitterDev e;
  std::memset(&e, 0, sizeof(e));
  e.width = width;
  e.height = height;
  e.x = mk3(position[0], position[1], position[2]);
  for (int k = 0; k < 9; k++) e.rot.a[k] = rot[k];
  e.velocity = velocity;
  e.emit_start = emit_start;
  e.emit_end = emit_end;
  e.next_emit_time = emit_start;
  e.emit_counter = 0;
  c->h_emitters.push_back(e);
  return DFR_OK;
}

// Host-only: cut nz cell layers into n_ranks ranges holding (nearly) equal numbers of particles.
int dfr_slab_plan(double z_origin, double inv_cell, int nz, int reach, int64_t n, const double *z, int n_ranks, int32_t *planes) {
  if (nz <= 0 || n_ranks < 1 || !planes || (n > 0 && !z)) return DFR_ERR_INVALID;
  std::vector<int64_t> hist(nz, 0);
  for (int64_t i = 0; i < n; i++) {
    int zc = (int)std::floor((z[i] - z_origin) * inv_cell);  // the arithmetic of cell_of (dfr_kernels.cuh), global origin
    zc = std::min(std::max(zc, 0), nz - 1);
    hist[zc]++;
  }
  planes[0] = 0;
  planes[n_ranks] = nz;
  int64_t cum = 0;
  int k = 1;
  for (int zc = 0; zc < nz && k < n_ranks; zc++) {
    cum += hist[zc];
    while (k < n_ranks && cum >= (n * (int64_t)k) / n_ranks) planes[k++] = zc + 1;
  }
  for (; k < n_ranks; k++) planes[k] = nz;
  for (int r = 0; r < n_ranks; r++)
    if (n_ranks > 1 && planes[r + 1] - planes[r] < 2 * reach + 1) return DFR_ERR_INVALID;
  return DFR_OK;
}

int dfr_slab_unique_id(char out[DFR_SLAB_ID_BYTES]) {
  std::string why;
  NcclApi *N = nccl_api(&why);
  if (!N || !out) return DFR_ERR_INVALID;
  static_assert(sizeof(ncclUniqueId) <= DFR_SLAB_ID_BYTES, "ncclUniqueId grew");
  ncclUniqueId id;
  if (N->GetUniqueId(&id) != ncclSuccess) return DFR_ERR_CUDA;
  std::memset(out, 0, DFR_SLAB_ID_BYTES);
  std::memcpy(out, &id, sizeof(id));
  return DFR_OK;
}

int dfr_slab_configure(dfr_context *c, int rank, int n_ranks, const char id_bytes[DFR_SLAB_ID_BYTES]) {
  if (!c) return DFR_ERR_INVALID;
  if (c->finalized) return fail(c, DFR_ERR_STATE, "slab_configure after finalize");
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks || !id_bytes) return fail(c, DFR_ERR_INVALID, "bad slab rank");
  if (n_ranks == 1) return DFR_OK;  // one slab is the plain context
  std::string why;
  NcclApi *N = nccl_api(&why);
  if (!N) return fail(c, DFR_ERR_INVALID, why);
  cudaSetDevice(c->device);
  ncclUniqueId id;
  std::memcpy(&id, id_bytes, sizeof(id));
  NC(N->CommInitRank(&c->slab.comm, n_ranks, id, rank));
  c->slab.on = true;
  c->slab.rank = rank;
  c->slab.n = n_ranks;
  return DFR_OK;
}

int dfr_slab_info(dfr_context *c, int64_t out[4]) {
  if (!c || !c->finalized || !out) return fail(c, DFR_ERR_STATE, "not finalized");
  int rc = sync_state(c);
  if (rc) return rc;
  out[0] = c->hSt->own_end - c->hSt->own_begin;
  out[1] = c->hSt->nf - out[0];
  out[2] = c->slab.exchanged_bytes;
  out[3] = c->slab.on ? (c->slab.p2p ? -c->slab.n : c->slab.n) : 1;  // negative: ghost updates travel as peer stores
  return DFR_OK;
}

int dfr_finalize(dfr_context *c) {
  if (!c) return DFR_ERR_INVALID;
  if (c->finalized) return fail(c, DFR_ERR_STATE, "already finalized");
  cudaSetDevice(c->device);
  set_cache_preferences();
  const dfr_config &cfg = c->cfg;
  Params &P = c->P;
  std::memset(&P, 0, sizeof(P));
  // ---- kernels & constants (Simulation.cpp:382-386, SPHKernels.h:25-34, FluidModel.cpp:255-275) ----
  const double hs = 4.0 * cfg.particle_radius;
  P.support_radius = hs;
  P.r2 = hs * hs;
  P.inv_h = 1.0 / hs;
  const double h3 = hs * hs * hs;
  P.k_cubic = 8.0 / (M_PI * h3);
  P.l_cubic = 48.0 / (M_PI * h3);
  P.W_zero = P.k_cubic;
  P.coh_k = 32.0 / (M_PI * std::pow(hs, 9.0));
  P.coh_c = std::pow(hs, 6.0) / 64.0;
  P.adh_k = 0.007 / std::pow(hs, 3.25);
  P.particle_radius = cfg.particle_radius;
  P.density0 = cfg.density0;
  const double diam = 2.0 * cfg.particle_radius;
  P.volume = 0.8 * diam * diam * diam;
  P.mass = P.volume * cfg.density0;
  P.gx = cfg.gravitation[0]; P.gy = cfg.gravitation[1]; P.gz = cfg.gravitation[2];
  P.cfl_factor = cfg.cfl_factor; P.cfl_min = cfg.cfl_min_time_step; P.cfl_max = cfg.cfl_max_time_step;
  P.cfl_method = cfg.cfl_method;
  P.min_iter = cfg.min_iterations; P.max_iter = cfg.max_iterations; P.max_iter_v = cfg.max_iterations_v;
  P.max_error = cfg.max_error; P.max_error_v = cfg.max_error_v;
  P.use_warm_p = cfg.use_pressure_warmstart; P.use_warm_v = cfg.use_divergence_warmstart;
  P.visc_method = cfg.viscosity_method; P.st_method = cfg.surface_tension_method;
  P.viscosity = cfg.viscosity; P.viscosity_b = cfg.viscosity_boundary;
  P.surface_tension = cfg.surface_tension; P.surface_tension_b = cfg.surface_tension_boundary;
  P.gradient_mode = cfg.gradient_mode; P.rigid_body_mode = cfg.rigid_body_mode; P.optimize_rotation = cfg.optimize_rotation;
  P.use_manager = cfg.use_rigid_gradient_manager; P.use_contact = cfg.use_rigid_contact_solver;
  P.target_time = cfg.target_time; P.uniform_acc_time = cfg.uniform_acc_rb_time;
  P.release_mode = cfg.use_release_rigid_body_mode != 0;
  P.time_step_size0 = cfg.time_step_size;
  P.n_bodies = (int)c->bodies.size();
  // ---- boundary layout: static bodies first, then dynamic ----
  int off = 0;
  std::vector<int> order;
  for (size_t i = 0; i < c->bodies.size(); i++)
    if (!c->bodies[i].dynamic) order.push_back((int)i);
  c->n_static_p = 0;
  for (int i : order) {
    c->bodies[i].p_begin = off;
    off += (int)c->bodies[i].n;
  }
  c->n_static_p = off;
  c->dyn_begin = off;
  int ndynb = 0;
  for (size_t i = 0; i < c->bodies.size(); i++)
    if (c->bodies[i].dynamic) {
      c->bodies[i].p_begin = off;
      off += (int)c->bodies[i].n;
      order.push_back((int)i);
      ndynb++;
    }
  c->n_b = off;
  c->n_dyn_p = c->n_b - c->n_static_p;
  P.n_dyn_bodies = ndynb;

  // ---- world positions on the host (for the bounding box) ----
  std::vector<double4> h_bpos(c->n_b), h_bx0(c->n_b);
  std::vector<int> h_bbody(c->n_b), h_borig(c->n_b);
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  auto grow = [&](double x, double y, double z) {
    lo[0] = std::min(lo[0], x); lo[1] = std::min(lo[1], y); lo[2] = std::min(lo[2], z);
    hi[0] = std::max(hi[0], x); hi[1] = std::max(hi[1], y); hi[2] = std::max(hi[2], z);
  };
  for (size_t bi = 0; bi < c->bodies.size(); bi++) {
    HostBody &hb = c->bodies[bi];
    host_body_record(c, hb, hb.p_begin);
    const m33 R = qrot(hb.dev0.q);
    for (int64_t j = 0; j < hb.n; j++) {
      const d3 x0 = mk3(hb.x_local[3 * j], hb.x_local[3 * j + 1], hb.x_local[3 * j + 2]);
      const d3 x = R * x0 + hb.dev0.pos;
      const int g = hb.p_begin + (int)j;
      h_bpos[g] = make_double4(x.x, x.y, x.z, 0.0);
      h_bx0[g] = make_double4(x0.x, x0.y, x0.z, 0.0);
      h_bbody[g] = (int)bi;
      h_borig[g] = g;
      grow(x.x, x.y, x.z);
    }
  }
  for (int64_t i = 0; i < c->nf0; i++) grow(c->h_fx[3 * i], c->h_fx[3 * i + 1], c->h_fx[3 * i + 2]);
  if (c->nf0 == 0 && c->n_b == 0) return fail(c, DFR_ERR_INVALID, "empty scene");
  // ---- grid: cell edge slightly above (support radius / reach), margin cells, capped cell count ----
  int reach = c->cfg.grid_reach > 0 ? c->cfg.grid_reach : 2;
  double cell = hs * (1.0 + 1.0e-7) / reach;
  double margin = 2.0 * reach;
  for (;;) {
    double nx = std::floor((hi[0] - lo[0]) / cell) + 1 + 2 * margin;
    double ny = std::floor((hi[1] - lo[1]) / cell) + 1 + 2 * margin;
    double nz = std::floor((hi[2] - lo[2]) / cell) + 1 + 2 * margin;
    if (nx * ny * nz <= (double)(1 << 26)) {
      P.grid.nx = (int)nx; P.grid.ny = (int)ny; P.grid.nz = (int)nz;
      break;
    }
    if (reach > 1) {  // fall back to support-radius cells before growing them
      reach = 1;
      cell = hs * (1.0 + 1.0e-7);
      margin = 2.0;
    } else
      cell *= 1.26;
  }
  P.grid.reach = reach;
  P.grid.ox = lo[0] - margin * cell; P.grid.oy = lo[1] - margin * cell; P.grid.oz = lo[2] - margin * cell;
  P.grid.inv_cell = 1.0 / cell;
  P.grid.ncells = P.grid.nx * P.grid.ny * P.grid.nz;

  // ---- slab decomposition: cut the z layers into ranges of equal particle count, keep my range (+ pad) ----
  c->nf_loc0 = c->nf0;
  int slab_ghost_estimate = 0;
  P.n_global = c->nf0;
  if (c->slab.on) {
    if (!c->h_emitters.empty()) return fail(c, DFR_ERR_INVALID, "slab decomposition: emitters are not supported");
    if (cfg.use_rigid_contact_solver) return fail(c, DFR_ERR_INVALID, "slab decomposition: the rigid contact solver is not supported");
    auto &S = c->slab;
    GridGeom &G = P.grid;
    std::vector<int64_t> hist(G.nz, 0);
    std::vector<int> zc(c->nf0);
    for (int64_t i = 0; i < c->nf0; i++) {
      int z = (int)std::floor((c->h_fx[3 * i + 2] - G.oz) * G.inv_cell);
      z = std::min(std::max(z, 0), G.nz - 1);
      zc[i] = z;
      hist[z]++;
    }
    std::vector<int> planes(S.n + 1, 0);
    {
      std::vector<double> zs(c->nf0);
      for (int64_t i = 0; i < c->nf0; i++) zs[i] = c->h_fx[3 * i + 2];
      const int prc = dfr_slab_plan(G.oz, G.inv_cell, G.nz, G.reach, c->nf0, zs.data(), S.n, planes.data());
      if (prc == DFR_ERR_INVALID)
        return fail(c, DFR_ERR_INVALID, "slab decomposition: a slab would be thinner than two support radii (too many ranks for this scene)");
    }
    const int zlo = planes[S.rank], zhi = planes[S.rank + 1];
    const int pad = 2 * G.reach + 2;
    S.G.reach = G.reach;
    S.G.has_lo = S.rank > 0;
    S.G.has_hi = S.rank < S.n - 1;
    S.G.own_zlo = pad;
    S.G.own_zhi = pad + (zhi - zlo);
    // particles per exchanged layer set (reach + 1 layers on either side of either plane), for the buffer sizes
    int64_t est = 0;
    auto layers = [&](int z0, int z1) {
      int64_t s2 = 0;
      for (int z = std::max(z0, 0); z < std::min(z1, G.nz); z++) s2 += hist[z];
      est = std::max(est, s2);
    };
    layers(zlo - G.reach - 1, zlo);
    layers(zlo, zlo + G.reach + 1);
    layers(zhi - G.reach - 1, zhi);
    layers(zhi, zhi + G.reach + 1);
    slab_ghost_estimate = (int)est;
    G.z_shift = zlo - pad;
    G.nz = (zhi - zlo) + 2 * pad;
    G.ncells = G.nx * G.ny * G.nz;
    c->h_ids0.clear();
    for (int64_t i = 0; i < c->nf0; i++)
      if (zc[i] >= zlo && zc[i] < zhi) c->h_ids0.push_back((int)i);
    c->nf_loc0 = (int64_t)c->h_ids0.size();
    P.slab = 1;
  }

  // ---- allocations ----
  c->nf_cap = c->nf0 + std::max(0, cfg.max_emitted_particles);
  if (c->slab.on) {
    // room for the flow to pile up in my slab (x1.5) and for two ghost layers
    c->nf_cap = c->nf_loc0 + c->nf_loc0 / 2 + 4 * (int64_t)slab_ghost_estimate + 16384;
    c->slab.send_cap = 3 * slab_ghost_estimate + 8192;
  }
  const size_t N = (size_t)std::max<int64_t>(c->nf_cap, 1);
  c->launch_nf = c->h_emitters.empty() ? (int)c->nf_loc0 : (int)c->nf_cap;  // emitters grow st->nf on the device
  if (c->slab.on) {
    auto &S = c->slab;
    for (int k = 0; k < 2; k++) {
      CU(S.s_pos[k].alloc(S.send_cap)); CU(S.s_vel[k].alloc(S.send_cap)); CU(S.s_misc[k].alloc(S.send_cap));
    }
    CU(S.r_misc.alloc(2 * (size_t)S.send_cap));
    CU(S.counts.alloc(8));
    std::memset(&S.xchg, 0, sizeof(S.xchg));  // the receive area is sized in slab_p2p_setup (one capacity for all ranks)
    S.xchg_ok = false;
    CU(S.body_buf.alloc(std::max<size_t>(c->bodies.size(), 1) * ACC_N));
    if (cudaMallocHost((void **)&S.h_counts, 8 * sizeof(int)) != cudaSuccess) return fail(c, DFR_ERR_CUDA, "cudaMallocHost");
  }
  c->slab_needs_p2p_setup = c->slab.on;
  if (c->slab.on && c->nf_loc0 > 0 && c->slab.h_stage_n < (size_t)c->nf_loc0) {
    // pinned staging of dfr_load_fluid_state, allocated here so that a load never pays for the (slow) pinned allocation
    if (c->slab.h_stage) cudaFreeHost(c->slab.h_stage);
    c->slab.h_stage = nullptr;
    c->slab.h_stage_n = 0;
    if (cudaMallocHost((void **)&c->slab.h_stage, (size_t)c->nf_loc0 * 10 * sizeof(double)) != cudaSuccess)
      return fail(c, DFR_ERR_CUDA, "cudaMallocHost");
    c->slab.h_stage_n = (size_t)c->nf_loc0;
  }
  CU(c->dEmitters.alloc(std::max<size_t>(c->h_emitters.size(), 1)));
  const int nc = P.grid.ncells;
  CU(c->dSt.alloc(1));
  CU(c->sched_ctr.alloc(2 * DFR_SCHED_STRIDE));
  {
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, c->device));
    c->nsm = std::max(1, std::min(prop.multiProcessorCount, (int)DFR_SCHED_STRIDE));
  }
  for (int k = 0; k < 2; k++) {
    CU(c->pos[k].alloc(N)); CU(c->vel[k].alloc(N)); CU(c->kappa[k].alloc(N)); CU(c->kappav[k].alloc(N));
    CU(c->pid[k].alloc(N)); CU(c->pstate[k].alloc(N));
  }
  CU(c->acc.alloc(N)); CU(c->sgp.alloc(N)); CU(c->normal.alloc(N)); CU(c->density.alloc(N)); CU(c->factor.alloc(N));
  CU(c->dadv.alloc(N)); CU(c->xk.alloc(N)); CU(c->xrho.alloc(N)); CU(c->partials.alloc((N / 128 + 2) * 4 + RES_BLOCKS));  // per-warp residual partials + the slice sums of k_residual_finish
  CU(c->pos_init.alloc(N)); CU(c->vel_init.alloc(N)); CU(c->kappa_init.alloc(N)); CU(c->kappav_init.alloc(N));
  const size_t NB = (size_t)std::max(c->n_b, 1);
  CU(c->bpos.alloc(NB)); CU(c->bvel.alloc(NB)); CU(c->bx0.alloc(NB)); CU(c->bbody.alloc(NB)); CU(c->borig.alloc(NB));
  CU(c->bvol.alloc(NB));
  CU(c->dBodies.alloc(std::max<size_t>(c->bodies.size(), 1)));
  CU(c->dMgr.alloc(std::max<size_t>(c->bodies.size() * c->bodies.size(), 1)));
  if (cudaMallocHost((void **)&c->h_bodies, std::max<size_t>(c->bodies.size(), 1) * sizeof(BodyDev)) != cudaSuccess) return fail(c, DFR_ERR_CUDA, "cudaMallocHost");
  if (cudaMallocHost((void **)&c->h_init, std::max<size_t>(c->bodies.size(), 1) * 6 * sizeof(double)) != cudaSuccess) return fail(c, DFR_ERR_CUDA, "cudaMallocHost");
  CU(c->cell_start_f.alloc((size_t)nc + 1)); CU(c->cell_start_s.alloc((size_t)nc + 1)); CU(c->cell_start_d.alloc((size_t)nc + 1));
  CU(c->near_s.alloc((size_t)nc));
  CU(c->near_d.alloc((size_t)nc));
  CU(cudaMemsetAsync(c->near_s.p, 0, (size_t)nc, c->stream));
  const size_t max_scan = std::max<size_t>((size_t)nc + 1, (size_t)c->n_dyn_p + 1);
  CU(c->tile_sums.alloc(max_scan / SCAN_TILE + 2));
  CU(c->tile_sums_side.alloc(max_scan / SCAN_TILE + 2));
  CU(c->cell_of_p.alloc(N)); CU(c->rank_in_cell.alloc(N)); CU(c->sorted_src_f.alloc(N));
  const size_t ND = (size_t)std::max(c->n_dyn_p, 1), NS = (size_t)std::max(c->n_static_p, 1);
  CU(c->sorted_src_d.alloc(std::max(ND, NS))); CU(c->cell_of_b.alloc(std::max(ND, NS))); CU(c->rank_b.alloc(std::max(ND, NS)));
  CU(c->cnt_f.alloc(N)); CU(c->cnt_b.alloc(N));
  // ELL capacities: neighbours per particle (support 4r, spacing 2r: ~30 at rest, more under compression).  Slots
  // beyond a row's count are never touched, so generous capacities cost address space, not bandwidth.
  c->cap_f = ((cfg.neighbor_capacity_fluid > 0 ? cfg.neighbor_capacity_fluid : 96) + 3) & ~3;
  c->cap_b = ((cfg.neighbor_capacity_boundary > 0 ? cfg.neighbor_capacity_boundary : 64) + 3) & ~3;
  const int per_d = cfg.body_neighbor_capacity > 0 ? cfg.body_neighbor_capacity : 96;
  c->cap_d = (unsigned int)(ND * per_d);
  const size_t nwarp = (N + 31) / 32;
  CU(c->idx_f.alloc(nwarp * 32 * (size_t)c->cap_f)); CU(c->idx_b.alloc(nwarp * 32 * (size_t)c->cap_b));
  CU(c->idx_d.alloc(c->cap_d)); CU(c->off_d.alloc(ND + 1));
  if (c->slab_needs_p2p_setup) {  // needs the gathered arrays to exist
    int prc = slab_p2p_setup(c);
    if (prc) return prc;
    c->slab_needs_p2p_setup = false;
  }

  // ---- accumulator blocks of the boundary-side kernel: one body per block ----
  std::vector<int> blk_body, blk_first;
  for (size_t bi = 0; bi < c->bodies.size(); bi++) {
    HostBody &hb = c->bodies[bi];
    hb.blk_begin = (int)blk_body.size();
    hb.blk_count = 0;
    if (!hb.dynamic) continue;
    for (int f = 0; f < (int)hb.n; f += BS_PART_PER_BLOCK) {
      blk_body.push_back((int)bi);
      blk_first.push_back(hb.p_begin + f);
      hb.blk_count++;
    }
  }
  c->n_acc_blocks = (int)blk_body.size();
  CU(c->blk_body.alloc(std::max<size_t>(blk_body.size(), 1)));
  CU(c->blk_first.alloc(std::max<size_t>(blk_first.size(), 1)));
  CU(c->acc_rows.alloc(std::max<size_t>(blk_body.size(), 1) * ACC_N));
  if (!blk_body.empty()) {
    CU(cudaMemcpy(c->blk_body.p, blk_body.data(), blk_body.size() * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->blk_first.p, blk_first.data(), blk_first.size() * sizeof(int), cudaMemcpyHostToDevice));
  }

  // ---- uploads ----
  if (c->nf_loc0) {
    std::vector<double4> p4(c->nf_loc0), v4(c->nf_loc0);
    for (int64_t k = 0; k < c->nf_loc0; k++) {
      const int64_t i = c->slab.on ? c->h_ids0[k] : k;
      p4[k] = make_double4(c->h_fx[3 * i], c->h_fx[3 * i + 1], c->h_fx[3 * i + 2], 0.0);
      v4[k] = make_double4(c->h_fv[3 * i], c->h_fv[3 * i + 1], c->h_fv[3 * i + 2], 0.0);
    }
    CU(cudaMemcpy(c->pos_init.p, p4.data(), c->nf_loc0 * sizeof(double4), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->vel_init.p, v4.data(), c->nf_loc0 * sizeof(double4), cudaMemcpyHostToDevice));
  }
  if (c->n_b) {
    CU(cudaMemcpy(c->bpos.p, h_bpos.data(), c->n_b * sizeof(double4), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->bx0.p, h_bx0.data(), c->n_b * sizeof(double4), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->bbody.p, h_bbody.data(), c->n_b * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->borig.p, h_borig.data(), c->n_b * sizeof(int), cudaMemcpyHostToDevice));
  }
  c->h_borig = h_borig;
  // ---- one-time physical sort of the static boundary particles into cell order ----
  if (c->n_static_p > 0) {
    const int ns = c->n_static_p;
    CU(c->bpos_tmp.alloc(ns)); CU(c->bx0_tmp.alloc(ns)); CU(c->bbody_tmp.alloc(ns)); CU(c->borig_tmp.alloc(ns));
    CU(cudaMemsetAsync(c->cell_start_s.p, 0, sizeof(unsigned int) * (nc + 1), c->stream));
    LAUNCH(c, k_bin_count, cdiv(ns, 128), 128, c->P, c->bpos.p, (const int *)nullptr, ns, c->cell_start_s.p, c->cell_of_b.p,
           c->rank_b.p);
    int rc = scan_u32(c, c->cell_start_s.p, (size_t)nc + 1, nullptr);
    if (rc) return rc;
    LAUNCH(c, k_bin_scatter, cdiv(ns, 128), 128, (const int *)nullptr, ns, c->cell_start_s.p, c->cell_of_b.p, c->rank_b.p,
           c->sorted_src_d.p);
    LAUNCH(c, k_bin_sort_cells, cdiv(nc, 128), 128, c->cell_start_s.p, nc, c->sorted_src_d.p);
    LAUNCH(c, k_permute_boundary, cdiv(ns, 128), 128, ns, c->sorted_src_d.p, c->bpos.p, c->bx0.p, c->bbody.p, c->borig.p,
           c->bpos_tmp.p, c->bx0_tmp.p, c->bbody_tmp.p, c->borig_tmp.p);
    CU(cudaMemcpyAsync(c->bpos.p, c->bpos_tmp.p, ns * sizeof(double4), cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(c->bx0.p, c->bx0_tmp.p, ns * sizeof(double4), cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(c->bbody.p, c->bbody_tmp.p, ns * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(c->borig.p, c->borig_tmp.p, ns * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(c->h_borig.data(), c->borig.p, ns * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    LAUNCH(c, k_mark_near, cdiv(ns, 128), 128, c->P, c->bpos.p, ns, c->near_s.p);
    CU(cudaStreamSynchronize(c->stream));
    c->bpos_tmp.free(); c->bx0_tmp.free(); c->bbody_tmp.free(); c->borig_tmp.free();
  }
  c->finalized = true;
  int rc = reset_device_state(c);
  if (rc) {
    c->finalized = false;
    return rc;
  }
  CU(cudaGetLastError());
  return DFR_OK;
}

int64_t dfr_slab_local_ids(dfr_context *c, int32_t *ids_out, int64_t capacity) {
  if (!c || !c->finalized) return 0;
  const int64_t n = c->nf_loc0;
  if (ids_out)
    for (int64_t k = 0; k < n && k < capacity; k++) ids_out[k] = c->slab.on ? c->h_ids0[k] : (int32_t)k;
  return n;
}

static int load_fluid_state_impl(dfr_context *c, const double *x, const double *v, const double *kappa, const double *kappa_v, bool local_rows);
int dfr_load_fluid_state_local(dfr_context *c, int64_t n, const double *x, const double *v, const double *kappa, const double *kappa_v) {
  if (!c || !c->finalized) return fail(c, DFR_ERR_STATE, "not finalized");
  if (n != c->nf_loc0) return fail(c, DFR_ERR_INVALID, "dfr_load_fluid_state_local: n differs from the rows this context holds (dfr_slab_local_ids)");
  return load_fluid_state_impl(c, x, v, kappa, kappa_v, true);
}
int dfr_load_fluid_state(dfr_context *c, const double *x, const double *v, const double *kappa, const double *kappa_v) {
  return load_fluid_state_impl(c, x, v, kappa, kappa_v, false);
}
static int load_fluid_state_impl(dfr_context *c, const double *x, const double *v, const double *kappa, const double *kappa_v, bool local_rows) {
  if (!c || !c->finalized) return fail(c, DFR_ERR_STATE, "not finalized");
  cudaSetDevice(c->device);
  const int64_t n = c->nf_loc0;  // arrays are indexed by particle id; a slab keeps the ids it owned at t = 0
  if (!c->slab.on || local_rows) {
    // straight from the caller's (ideally pinned) arrays: H2D into scratch, repack on the device
    double *stage = (double *)c->acc.p;  // n double4 of scratch >= 3 n doubles; reset clears it afterwards
    if (x && n) {
      CU(cudaMemcpyAsync(stage, x, 3 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      LAUNCH(c, k_xyz_to_rec, cdiv(n, 256), 256, stage, c->pos_init.p, (int)n);
    }
    if (v && n) {
      double *stage_v = (double *)c->sgp.p;
      CU(cudaMemcpyAsync(stage_v, v, 3 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      LAUNCH(c, k_xyz_to_rec, cdiv(n, 256), 256, stage_v, c->vel_init.p, (int)n);
    }
    if (kappa && n) CU(cudaMemcpyAsync(c->kappa_init.p, kappa, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (kappa_v && n) CU(cudaMemcpyAsync(c->kappav_init.p, kappa_v, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return reset_device_state(c);
  }
  // slab: this rank keeps the ids it owned at t = 0.  The rows are gathered by a few host threads into a pinned staging
  // buffer and copied asynchronously (a single-threaded gather into pageable vectors + blocking copies cost ~35 ms per
  // million particles, i.e. most of a short trajectory's end-to-end time)
  if (n > 0) {
    if (c->slab.h_stage_n < (size_t)n) {
      if (c->slab.h_stage) cudaFreeHost(c->slab.h_stage);
      c->slab.h_stage = nullptr;
      if (cudaMallocHost((void **)&c->slab.h_stage, (size_t)n * 10 * sizeof(double)) != cudaSuccess) return fail(c, DFR_ERR_CUDA, "cudaMallocHost");
      c->slab.h_stage_n = (size_t)n;
    }
    double4 *sx = reinterpret_cast<double4 *>(c->slab.h_stage);  // n double4 | n double4 | n double | n double
    double4 *sv = sx + n;
    double *sk = reinterpret_cast<double *>(sv + n);
    double *skv = sk + n;
    const int *ids = c->h_ids0.data();
    const int nthreads = (int)std::max<int64_t>(1, std::min<int64_t>(8, n / 65536));
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; t++)
      pool.emplace_back([=]() {
        const int64_t lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
        for (int64_t k = lo; k < hi; k++) {
          const int64_t id = ids[k];
          if (x) sx[k] = make_double4(x[3 * id], x[3 * id + 1], x[3 * id + 2], 0.0);
          if (v) sv[k] = make_double4(v[3 * id], v[3 * id + 1], v[3 * id + 2], 0.0);
          if (kappa) sk[k] = kappa[id];
          if (kappa_v) skv[k] = kappa_v[id];
        }
      });
    for (auto &th : pool) th.join();
    if (x) CU(cudaMemcpyAsync(c->pos_init.p, sx, n * sizeof(double4), cudaMemcpyHostToDevice, c->stream));
    if (v) CU(cudaMemcpyAsync(c->vel_init.p, sv, n * sizeof(double4), cudaMemcpyHostToDevice, c->stream));
    if (kappa) CU(cudaMemcpyAsync(c->kappa_init.p, sk, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (kappa_v) CU(cudaMemcpyAsync(c->kappav_init.p, skv, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  // like the oracle, loading re-bases the running state: positions/velocities/kappas are replaced in id order
  return reset_device_state(c);
}

int dfr_reset(dfr_context *c) {
  if (!c || !c->finalized) return fail(c, DFR_ERR_STATE, "not finalized");
  cudaSetDevice(c->device);
  return reset_device_state(c);
}

int dfr_reset_gradient(dfr_context *c) {
  if (!c || !c->finalized) return fail(c, DFR_ERR_STATE, "not finalized");
  cudaSetDevice(c->device);
  c->bodies_mirrored = false;
  if (c->P.n_bodies > 0) {
    LAUNCH(c, k_reset_gradient, 1, 32, c->P, c->dBodies.p);
    if (c->acc_rows.n) CU(cudaMemsetAsync(c->acc_rows.p, 0, c->acc_rows.n * sizeof(double), c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  return DFR_OK;
}

int dfr_set_gradient_mode(dfr_context *c, int mode) {
  if (!c) return DFR_ERR_INVALID;
  if (mode < 0 || mode > 2) return fail(c, DFR_ERR_INVALID, "gradient mode must be 0 (Complete), 1 (Incomplete) or 2 (RigidGradOnly)");
  c->cfg.gradient_mode = mode;
  c->P.gradient_mode = mode;  // Params travel by value with every launch (__grid_constant__)
  drop_step_graphs(c);        // ... and were recorded into the step graphs
  return DFR_OK;
}

namespace {
// Enqueues n steps: recorded graphs where possible, the stream path otherwise.  gated = 1: steps after the end of the
// trajectory are skipped on the device (the first step is never skipped).  hSt must be current on entry (every public
// entry point that steps ends with a state read-back).
struct StepBatch {
  int64_t graph_steps = 0;
  const dfr_context::StepGraph *last = nullptr;
  long long it0 = 0, itv0 = 0;  // iteration totals when the first graph step of the batch was enqueued
};
// Graph stepping cannot look at the lists between the build and their first use, so it keeps the capacities at more
// than twice the longest row seen (a row does not double within one CFL-limited step; an overflow still surfaces as
// DFR_ERR_CAPACITY at the next read-back).  Unused ELL slots cost address space, not bandwidth.  The lists are rebuilt
// from scratch every step, so re-allocating between two steps loses nothing.
int relax_list_capacity(dfr_context *c) {
  const bool f = 2 * (int)c->hSt->list_used_f > c->cap_f, b = 2 * (int)c->hSt->list_used_b > c->cap_b;
  const bool d = 2ull * c->hSt->list_used_d > c->cap_d;
  if (!f && !b && !d) return DFR_OK;
  CU(cudaStreamSynchronize(c->stream));
  const size_t nwarp = ((size_t)c->nf_cap + 31) / 32;
  if (f) {
    const int want = ((int)(2 * c->hSt->list_used_f + c->hSt->list_used_f / 2) + 8 + 3) & ~3;  // 2.5x: room to creep up
    if (want > 4096) return DFR_OK;  // stay on the watched stream path
    c->idx_f.free();
    CU(c->idx_f.alloc(nwarp * 32 * (size_t)want));
    c->cap_f = want;
  }
  if (b) {
    const int want = ((int)(2 * c->hSt->list_used_b + c->hSt->list_used_b / 2) + 8 + 3) & ~3;
    if (want > 4096) return DFR_OK;
    c->idx_b.free();
    CU(c->idx_b.alloc(nwarp * 32 * (size_t)want));
    c->cap_b = want;
  }
  if (d) {
    const unsigned long long want = 3ull * c->hSt->list_used_d + 1024;
    if (want > (1ull << 31)) return DFR_OK;
    c->idx_d.free();
    CU(c->idx_d.alloc((size_t)want));
    c->cap_d = (unsigned int)want;
  }
  drop_step_graphs(c);  // the list pointers and capacities are recorded in the graphs
  if (getenv_int("DFR_DEBUG"))
    std::fprintf(stderr, "[dfr] step %d: list capacities -> f %d b %d d %u (longest rows f %u b %u, d entries %u)\n", c->hSt->step_count, c->cap_f,
                 c->cap_b, c->cap_d, c->hSt->list_used_f, c->hSt->list_used_b, c->hSt->list_used_d);
  return DFR_OK;
}
int enqueue_steps(dfr_context *c, int n_steps, int gated, StepBatch &B) {
  for (int s = 0; s < n_steps; s++) {
    if (c->slab.on) {
      // Every rank must take the same path in the same step (the two paths make different numbers of ghost-update
      // passes): the choice only depends on things all ranks share.  The head of the step - k_begin_step and the particle
      // exchange with its two host read-backs - stays on the stream; the rest is one graph replay.
      if (!graph_stepping_possible(c) || c->fresh_steps > 0) {
        if (c->fresh_steps > 0) c->fresh_steps--;
        int rc = launch_step(c);
        if (rc) return rc;
        continue;
      }
      {  // row capacities: local buffers, re-allocating and re-recording changes nothing the neighbours can see
        int rc = relax_list_capacity(c);
        if (rc) return rc;
      }
      if (c->slab.xchg_ok) {  // the whole step is one replay: particle exchange over peer memory, nothing read back
        dfr_context::StepGraph &sgx = c->step_graph[0][c->cur];
        if (!sgx.exec) {
          int rc = capture_step_graph(c, 0);
          if (rc) return rc;
        }
        if (B.graph_steps == 0) {
          B.it0 = c->hSt->total_iters;
          B.itv0 = c->hSt->total_iters_v;
        }
        CU(cudaGraphLaunch(sgx.exec, c->stream));
        c->cur = 1 - c->cur;
        c->slab.h_stale = true;
        B.graph_steps++;
        B.last = &sgx;
        continue;
      }
      LAUNCH(c, k_begin_step, 1, 32, c->P, c->dSt.p, c->dBodies.p, fuse_nonpressure_enabled(c) ? 1 : 0);
      int rc = slab_exchange_and_sort(c);  // leaves hSt current
      if (rc) return rc;
      dfr_context::StepGraph &sgs = c->step_graph[0][c->cur];
      if (!sgs.exec) {
        rc = capture_step_graph(c, 0);
        if (rc) return rc;  // a rank that fell back alone would hang the others: report instead
      }
      if (B.graph_steps == 0) {
        B.it0 = c->hSt->total_iters;
        B.itv0 = c->hSt->total_iters_v;
      }
      CU(cudaGraphLaunch(sgs.exec, c->stream));
      c->vcur = 1 - c->vcur;  // the body's velocity-buffer flip (k_nonpressure / k_apply_accel)
      B.graph_steps++;
      B.last = &sgs;
      continue;
    }
    if (graph_stepping_possible(c) && c->fresh_steps == 0) {
      int rc = relax_list_capacity(c);
      if (rc) return rc;
    }
    const bool risky = c->fresh_steps > 0 || 2 * (int)c->hSt->list_used_f > c->cap_f || 2 * (int)c->hSt->list_used_b > c->cap_b ||
                       2ull * c->hSt->list_used_d > c->cap_d;
    if (!graph_stepping_possible(c) || risky) {
      const int cap_f = c->cap_f, cap_b = c->cap_b;
      const unsigned int cap_d = c->cap_d;
      int rc = launch_step(c);  // ends its solves with a read-back: hSt is current again
      if (rc) return rc;
      if (cap_f != c->cap_f || cap_b != c->cap_b || cap_d != c->cap_d) drop_step_graphs(c);  // lists were re-allocated
      continue;
    }
    const int g = (gated && s > 0) ? 1 : 0;
    dfr_context::StepGraph &sg = c->step_graph[g][c->cur];
    if (!sg.exec) {
      if (getenv_int("DFR_DEBUG")) std::fprintf(stderr, "[dfr] step %d: recording step graph (gated %d, parity %d)\n", c->hSt->step_count, g, c->cur);
      int rc = capture_step_graph(c, g);
      if (rc) {  // keep working without graphs (the error text stays in last_error until the next failure)
        c->graph_broken = 1;
        s--;
        continue;
      }
    }
    int rc = contact_sort_tick(c);
    if (rc) return rc;
    if (B.graph_steps == 0) {
      B.it0 = c->hSt->total_iters;
      B.itv0 = c->hSt->total_iters_v;
    }
    CU(cudaGraphLaunch(sg.exec, c->stream));
    c->cur = 1 - c->cur;  // the re-sort of the step flipped the buffers (vel flips twice per step)
    B.graph_steps++;
    B.last = &sg;
  }
  return DFR_OK;
}
// after the read-back that follows a batch: kernels the replayed steps launched
void account_graph_launches(dfr_context *c, const StepBatch &B, int64_t steps_executed) {
  if (!B.last) return;
  c->launches += steps_executed * B.last->n_static + (c->hSt->total_iters_v - B.itv0) * B.last->n_div_body +
                 (c->hSt->total_iters - B.it0) * B.last->n_prs_body;
  if (c->slab.on && c->slab.h_stale) {  // device-side exchange: the ranges of the last replayed step, read back with the step state
    const int *r = c->hSt->slab_ranges;
    c->slab.sync_rows = (c->slab.G.has_lo ? (r[2] - r[0]) + r[0] : 0) + (c->slab.G.has_hi ? (r[1] - r[3]) + (r[4] - r[1]) : 0);
    c->slab.exchanged_bytes += c->slab.sync_rows * 96 * steps_executed;  // exported boundary layers + imported ghosts, 96 B each
  }
  if (c->slab.on)  // ghost updates of the replayed steps (32-byte rows; the row count of the last exchange stands for all of them)
    c->slab.exchanged_bytes += c->slab.sync_rows * 32 * (steps_executed * c->slab.cap_syncs_static +
                                                         (c->hSt->total_iters_v - B.itv0) * c->slab.cap_syncs_div +
                                                         (c->hSt->total_iters - B.it0) * c->slab.cap_syncs_prs);
}
}  // namespace

int dfr_step(dfr_context *c, int n_steps) {
  if (!c || !c->finalized) return fail(c, DFR_ERR_STATE, "not finalized");
  cudaSetDevice(c->device);
  c->bodies_mirrored = false;
  {
    int rc = upload_staged_init(c);
    if (rc) return rc;
  }
  CU(cudaEventRecord(c->ev0, c->stream));
  StepBatch B;
  {
    int rc = enqueue_steps(c, n_steps, 0, B);
    if (rc) return rc;
  }
  CU(cudaEventRecord(c->ev1, c->stream));
  if (!c->bodies.empty())
    CU(cudaMemcpyAsync(c->h_bodies, c->dBodies.p, c->bodies.size() * sizeof(BodyDev), cudaMemcpyDeviceToHost, c->stream));
  int rc = sync_state(c);  // the only synchronisation of a replayed dfr_step: status + the mirror the getters read
  if (rc) return rc;
  account_graph_launches(c, B, B.graph_steps);
  c->bodies_mirrored = !c->bodies.empty();
  if (c->profiling) prof_resolve(c, c->hSt->div_iters, c->hSt->prs_iters);
  CU(cudaGetLastError());
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  c->device_ms += ms;
  return DFR_OK;
}

int dfr_run_trajectory(dfr_context *c, int max_steps, int *steps_done) {
  if (!c || !c->finalized) return fail(c, DFR_ERR_STATE, "not finalized");
  cudaSetDevice(c->device);
  c->bodies_mirrored = false;
  {
    int rc = upload_staged_init(c);
    if (rc) return rc;
  }
  CU(cudaEventRecord(c->ev0, c->stream));
  int s = 0;
  // batches of steps between two looks at `finished`; replayed steps past the end of the trajectory are skipped on the
  // device (IF node), steps on the stream path are enqueued one at a time as before
  const int batch_max = (graph_stepping_possible(c) && !c->slab.on && getenv_int("DFR_TRAJECTORY_BATCH") >= 0)
                            ? std::max(1, getenv_int("DFR_TRAJECTORY_BATCH") ? getenv_int("DFR_TRAJECTORY_BATCH") : 16)
                            : 1;
  while (s < max_steps) {
    const bool risky = c->fresh_steps > 0;
    const int nb = risky ? 1 : std::min(batch_max, max_steps - s);
    const int count0 = c->hSt->step_count, cur0 = c->cur;
    StepBatch B;
    int rc = enqueue_steps(c, nb, 1, B);
    if (rc) return rc;
    rc = sync_state(c);
    if (rc) return rc;
    const int done = c->hSt->step_count - count0;
    account_graph_launches(c, B, std::max(0, done - (nb - (int)B.graph_steps)));
    if (B.graph_steps) c->cur = cur0 ^ (done & 1);  // skipped steps did not flip the buffers
    s += done;
    if (c->hSt->finished || done < nb) break;
  }
  CU(cudaEventRecord(c->ev1, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (c->profiling) prof_resolve(c, c->hSt->div_iters, c->hSt->prs_iters);
  CU(cudaGetLastError());
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  c->device_ms += ms;
  if (steps_done) *steps_done = s;
  return DFR_OK;
}

int dfr_get_step_info(dfr_context *c, dfr_step_info *info) {
  if (!c || !info || !c->finalized) return fail(c, DFR_ERR_STATE, "not finalized");
  cudaSetDevice(c->device);
  int rc = sync_state(c);
  if (rc) return rc;
  const StepState &s = *c->hSt;
  info->time = s.time;
  info->time_step_size = s.h;
  info->iterations = s.last_iters;
  info->iterations_v = s.last_iters_v;
  info->step_count = s.step_count;
  info->trajectory_finished = s.finished;
  info->num_fluid_particles = c->slab.on ? (int64_t)(s.own_end - s.own_begin) : (int64_t)s.nf;  // slab mode: owned by this rank
  info->total_pressure_iterations = s.total_iters;
  info->total_divergence_iterations = s.total_iters_v;
  info->total_particle_steps = s.total_particle_steps;
  info->total_fluid_neighbors = s.total_neighbors;
  return DFR_OK;
}

// The getters below read the host mirror: one D2H of all body records per step (riding on the step's own state
// read-back) instead of one blocking copy per getter call - the scripts read 9 blocks per body and step.
static int fetch_body(dfr_context *c, int body, BodyDev &B) {
  if (!c || !c->finalized) return fail(c, DFR_ERR_STATE, "not finalized");
  if (body < 0 || body >= (int)c->bodies.size()) return fail(c, DFR_ERR_INVALID, "bad body index");
  if (!c->bodies_mirrored) {
    cudaSetDevice(c->device);
    CU(cudaMemcpyAsync(c->h_bodies, c->dBodies.p, c->bodies.size() * sizeof(BodyDev), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->bodies_mirrored = true;
  }
  B = c->h_bodies[body];
  return DFR_OK;
}

int dfr_get_body_state(dfr_context *c, int body, double out[13]) {
  BodyDev B;
  int rc = fetch_body(c, body, B);
  if (rc) return rc;
  out[0] = B.pos.x; out[1] = B.pos.y; out[2] = B.pos.z;
  out[3] = B.q.w; out[4] = B.q.x; out[5] = B.q.y; out[6] = B.q.z;
  out[7] = B.vel.x; out[8] = B.vel.y; out[9] = B.vel.z;
  out[10] = B.omega.x; out[11] = B.omega.y; out[12] = B.omega.z;
  return DFR_OK;
}

int dfr_set_body_velocity(dfr_context *c, int body, const double v[3], const double omega[3]) {
  if (!c || !c->finalized) return fail(c, DFR_ERR_STATE, "not finalized");
  if (body < 0 || body >= (int)c->bodies.size()) return fail(c, DFR_ERR_INVALID, "bad body index");
  cudaSetDevice(c->device);
  c->bodies_mirrored = false;
  BodyDev *d = c->dBodies.p + body;
  if (v) {
    const d3 t = mk3(v[0], v[1], v[2]);
    CU(cudaMemcpyAsync(&d->vel, &t, sizeof(d3), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  if (omega) {
    const d3 t = mk3(omega[0], omega[1], omega[2]);
    CU(cudaMemcpyAsync(&d->omega, &t, sizeof(d3), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  return DFR_OK;
}

int dfr_get_body_properties(dfr_context *c, int body, double out[17]) {
  BodyDev B;
  int rc = fetch_body(c, body, B);
  if (rc) return rc;
  out[0] = B.mass;
  out[1] = B.inv_mass;
  for (int k = 0; k < 9; k++) out[2 + k] = B.I0.a[k];
  out[11] = B.force_last.x; out[12] = B.force_last.y; out[13] = B.force_last.z;
  out[14] = B.torque_last.x; out[15] = B.torque_last.y; out[16] = B.torque_last.z;
  return DFR_OK;
}

int dfr_get_body_grad(dfr_context *c, int body, int which, double out[12]) {
  BodyDev B;
  int rc = fetch_body(c, body, B);
  if (rc) return rc;
  std::memset(out, 0, 12 * sizeof(double));
  switch (which) {
    case 0: put(out, B.x_v0); break;
    case 1: put(out, B.x_w0); break;
    case 2: put(out, B.q_v0); break;
    case 3: put(out, B.q_w0); break;
    case 4: put(out, B.v_v0); break;
    case 5: put(out, B.v_w0); break;
    case 6: put(out, B.w_v0); break;
    case 7: put(out, B.w_w0); break;
    case 8: put(out, B.net_f_v); break;
    case 9: put(out, B.net_f_x); break;
    case 10: put(out, B.net_f_q); break;
    case 11: put(out, B.net_f_w); break;
    case 12: put(out, B.net_t_v); break;
    case 13: put(out, B.net_t_x); break;
    case 14: put(out, B.net_t_q); break;
    case 15: put(out, B.net_t_w); break;
    default: return fail(c, DFR_ERR_INVALID, "bad gradient selector");
  }
  return DFR_OK;
}

int dfr_get_manager_grad(dfr_context *c, int R, int RR, int which, double out[12]) {
  if (!c || !c->finalized) return fail(c, DFR_ERR_STATE, "not finalized");
  const int n = (int)c->bodies.size();
  if (R < 0 || RR < 0 || R >= n || RR >= n) return fail(c, DFR_ERR_INVALID, "bad body index");
  cudaSetDevice(c->device);
  MgrBlock M;
  CU(cudaMemcpyAsync(&M, c->dMgr.p + (R * n + RR), sizeof(MgrBlock), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  std::memset(out, 0, 12 * sizeof(double));
  switch (which) {
    case 0: put(out, M.xn_v0); break;
    case 1: put(out, M.xn_w0); break;
    case 2: put(out, M.qn_v0); break;
    case 3: put(out, M.qn_w0); break;
    case 4: put(out, M.vn_v0); break;
    case 5: put(out, M.vn_w0); break;
    case 6: put(out, M.wn_v0); break;
    case 7: put(out, M.wn_w0); break;
    case 8: put(out, M.f_vn); break;
    case 9: put(out, M.f_xn); break;
    case 10: put(out, M.f_qn); break;
    case 11: put(out, M.f_wn); break;
    case 12: put(out, M.t_vn); break;
    case 13: put(out, M.t_xn); break;
    case 14: put(out, M.t_qn); break;
    case 15: put(out, M.t_wn); break;
    default: return fail(c, DFR_ERR_INVALID, "bad gradient selector");
  }
  return DFR_OK;
}

int64_t dfr_num_fluid(dfr_context *c) {
  if (!c || !c->finalized) return c ? c->nf0 : 0;
  cudaSetDevice(c->device);
  if (sync_state(c)) return 0;
  return c->slab.on ? c->nf0 : c->hSt->nf;  // arrays of the parity dumps are indexed by the scene's particle ids
}
int64_t dfr_num_fluid_initial(dfr_context *c) { return c ? c->nf0 : 0; }
int64_t dfr_num_body_particles(dfr_context *c, int body) {
  return (c && body >= 0 && body < (int)c->bodies.size()) ? c->bodies[body].n : 0;
}
int dfr_num_bodies(dfr_context *c) { return c ? (int)c->bodies.size() : 0; }

int dfr_download_fluid(dfr_context *c, int field, double *out) {
  if (!c || !c->finalized || !out) return fail(c, DFR_ERR_STATE, "not finalized");
  cudaSetDevice(c->device);
  int rc = sync_state(c);
  if (rc) return rc;
  // slab mode: only the particles this rank owns are written (entries of other ids stay untouched; the caller merges)
  const int i0 = c->slab.on ? c->hSt->own_begin : 0;
  const int n = c->slab.on ? c->hSt->own_end - i0 : c->hSt->nf;
  if (n <= 0) return DFR_OK;
  std::vector<int> ids(n);
  CU(cudaMemcpy(ids.data(), c->pid[c->cur].p + i0, n * sizeof(int), cudaMemcpyDeviceToHost));
  const double4 *v4 = nullptr;
  const double *s1 = nullptr;
  switch (field) {
    case 0: v4 = c->pos[c->cur].p; break;
    case 1: v4 = c->vel[c->vcur].p; break;
    case 2: s1 = c->density.p; break;
    case 3: s1 = c->factor.p; break;
    case 4: s1 = c->kappa[c->cur].p; break;
    case 5: s1 = c->kappav[c->cur].p; break;
    case 6: s1 = c->dadv.p; break;
    case 7: v4 = c->acc.p; break;
    case 8: v4 = c->sgp.p; break;
    case 9: v4 = c->normal.p; break;
    default: return fail(c, DFR_ERR_INVALID, "bad field");
  }
  if (v4) {
    std::vector<double4> t(n);
    CU(cudaMemcpy(t.data(), v4 + i0, n * sizeof(double4), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++) {
      double *o = out + 3 * (size_t)ids[i];
      o[0] = t[i].x; o[1] = t[i].y; o[2] = t[i].z;
    }
  } else {
    std::vector<double> t(n);
    CU(cudaMemcpy(t.data(), s1 + i0, n * sizeof(double), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++) out[ids[i]] = t[i];
  }
  return DFR_OK;
}

int dfr_download_body(dfr_context *c, int body, int field, double *out) {
  if (!c || !c->finalized || !out) return fail(c, DFR_ERR_STATE, "not finalized");
  if (body < 0 || body >= (int)c->bodies.size()) return fail(c, DFR_ERR_INVALID, "bad body index");
  cudaSetDevice(c->device);
  CU(cudaStreamSynchronize(c->stream));
  const HostBody &hb = c->bodies[body];
  const double4 *src = nullptr;
  switch (field) {
    case 0: case 2: src = c->bpos.p; break;
    case 1: src = c->bvel.p; break;
    case 3: src = c->bx0.p; break;
    default: return fail(c, DFR_ERR_INVALID, "bad field");
  }
  if (hb.dynamic) {
    std::vector<double4> t(hb.n);
    CU(cudaMemcpy(t.data(), src + hb.p_begin, hb.n * sizeof(double4), cudaMemcpyDeviceToHost));
    for (int64_t j = 0; j < hb.n; j++) {
      if (field == 2)
        out[j] = t[j].w;
      else {
        out[3 * j] = t[j].x; out[3 * j + 1] = t[j].y; out[3 * j + 2] = t[j].z;
      }
    }
  } else {
    std::vector<double4> t(c->n_static_p);
    CU(cudaMemcpy(t.data(), src, c->n_static_p * sizeof(double4), cudaMemcpyDeviceToHost));
    for (int s = 0; s < c->n_static_p; s++) {
      const int g = c->h_borig[s];
      if (g < hb.p_begin || g >= hb.p_begin + hb.n) continue;
      const int64_t j = g - hb.p_begin;
      if (field == 2)
        out[j] = t[s].w;
      else {
        out[3 * j] = t[s].x; out[3 * j + 1] = t[s].y; out[3 * j + 2] = t[s].z;
      }
    }
  }
  return DFR_OK;
}

int dfr_get_neighbors(dfr_context *c, int set_a, int set_b, int32_t *counts, int32_t *indices, int64_t capacity, int64_t *total) {
  if (!c || !c->finalized) return fail(c, DFR_ERR_STATE, "not finalized");
  const int nb = (int)c->bodies.size();
  if (set_a < -1 || set_a >= nb || set_b < -1 || set_b >= nb) return fail(c, DFR_ERR_INVALID, "bad set index");
  if (c->slab.on) return fail(c, DFR_ERR_INVALID, "neighbour dumps are not available on a slab-decomposed context");
  cudaSetDevice(c->device);
  // the same neighbourhood build the step runs, on the current positions
  int rc = build_neighbors(c);
  if (rc) return rc;
  c->fresh_steps = std::max(c->fresh_steps, 1);  // check the row capacities now
  rc = ensure_list_capacity(c);
  if (rc) return rc;
  rc = sync_state(c);
  if (rc) return rc;
  const int n = c->hSt->nf;
  std::vector<int> ids(std::max(n, 1));
  if (n) CU(cudaMemcpy(ids.data(), c->pid[c->cur].p, n * sizeof(int), cudaMemcpyDeviceToHost));
  std::vector<std::vector<int32_t>> rows;
  if (set_a == -1) {
    const bool fluid = (set_b == -1);
    rows.assign(n, {});
    if (n) {
      const size_t nw = ((size_t)n + 31) / 32;
      const size_t cap = fluid ? c->cap_f : c->cap_b;
      std::vector<int> cnt(n);
      CU(cudaMemcpy(cnt.data(), fluid ? c->cnt_f.p : c->cnt_b.p, n * sizeof(int), cudaMemcpyDeviceToHost));
      std::vector<int> idx(nw * 32 * cap);
      CU(cudaMemcpy(idx.data(), fluid ? c->idx_f.p : c->idx_b.p, idx.size() * sizeof(int), cudaMemcpyDeviceToHost));
      const HostBody *hb = fluid ? nullptr : &c->bodies[set_b];
      for (int i = 0; i < n; i++) {
        std::vector<int32_t> &r = rows[ids[i]];
        for (int k = 0; k < cnt[i]; k++) {
          const int j = idx[((((size_t)(i >> 5) * (cap >> 2) + (size_t)(k >> 2)) * 32 + (size_t)(i & 31)) << 2) + (size_t)(k & 3)];
          if (fluid)
            r.push_back(ids[j]);
          else {
            const int g = c->h_borig[j];
            if (g >= hb->p_begin && g < hb->p_begin + hb->n) r.push_back(g - hb->p_begin);
          }
        }
        std::sort(r.begin(), r.end());
      }
    }
  } else if (set_b == -1 && c->bodies[set_a].dynamic) {
    const HostBody &hb = c->bodies[set_a];
    rows.assign(hb.n, {});
    std::vector<unsigned int> off(c->n_dyn_p + 1);
    CU(cudaMemcpy(off.data(), c->off_d.p, (c->n_dyn_p + 1) * sizeof(unsigned int), cudaMemcpyDeviceToHost));
    std::vector<int> idx(std::max<size_t>(off[c->n_dyn_p], 1));
    if (off[c->n_dyn_p]) CU(cudaMemcpy(idx.data(), c->idx_d.p, (size_t)off[c->n_dyn_p] * sizeof(int), cudaMemcpyDeviceToHost));
    for (int64_t j = 0; j < hb.n; j++) {
      const int t = hb.p_begin - c->dyn_begin + (int)j;
      for (unsigned int p = off[t]; p < off[t + 1]; p++) rows[j].push_back(ids[idx[p]]);
      std::sort(rows[j].begin(), rows[j].end());
    }
  } else
    return fail(c, DFR_ERR_INVALID, "only fluid->fluid, fluid->body and dynamic body->fluid neighbour sets are stored (Simulation.cpp:882-900)");
  int64_t tot = 0;
  for (size_t i = 0; i < rows.size(); i++) {
    if (counts) counts[i] = (int32_t)rows[i].size();
    if (indices) {
      if (tot + (int64_t)rows[i].size() > capacity) return fail(c, DFR_ERR_CAPACITY, "neighbour buffer too small");
      std::memcpy(indices + tot, rows[i].data(), rows[i].size() * sizeof(int32_t));
    }
    tot += (int64_t)rows[i].size();
  }
  if (total) *total = tot;
  return DFR_OK;
}

int dfr_get_device_time_ms(dfr_context *c, double *total_ms, int64_t *kernel_launches) {
  if (!c) return DFR_ERR_INVALID;
  if (total_ms) *total_ms = c->device_ms;
  if (kernel_launches) *kernel_launches = c->launches;
  return DFR_OK;
}

int dfr_set_profiling(dfr_context *c, int enable) {
  if (!c) return DFR_ERR_INVALID;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->profiling) prof_resolve(c, 1 << 30, 1 << 30);
  c->profiling = enable != 0;
  c->prof_rows.clear();
  return DFR_OK;
}

int dfr_get_kernel_profile(dfr_context *c, int index, char *name, int name_capacity, double *total_ms, int64_t *launches) {
  if (!c) return DFR_ERR_INVALID;
  if (index < 0 || index >= (int)c->prof_rows.size()) return DFR_ERR_INVALID;
  auto it = c->prof_rows.begin();
  std::advance(it, index);
  if (name && name_capacity > 0) {
    std::strncpy(name, it->first.c_str(), (size_t)name_capacity - 1);
    name[name_capacity - 1] = 0;
  }
  if (total_ms) *total_ms = it->second.ms;
  if (launches) *launches = it->second.n;
  return DFR_OK;
}

}  // extern "C"
