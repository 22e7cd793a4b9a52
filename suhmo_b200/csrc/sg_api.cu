// sg_api.cu -- C-ABI implementation (include/suhmo_gpu.h): layouts, device-resident fields, the
// VCAMRNonLinearPoissonOp surface, the factory and the device-resident FAS multigrid driver.
// Host logic only; all arithmetic is in sg_kernels.cuh.  There is no CPU fallback anywhere in this file.
#include "../../include/suhmo_gpu.h"
#include <nvtx3/nvToolsExt.h> // header-only NVTX 3: ranges cost a predicted branch unless a profiler injected itself
#include <type_traits>
#include "sg_kernels.cuh"
#include "sg_twin.cuh"
#include "sg_general.cuh"
#include "sg_picard.cuh"
#include "sg_linear.cuh"
#include "sg_nccl.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
// NVTX range for the timeline views of Nsight Systems / Compute (V-cycle, level relaxations, operator update, residual norm)
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
extern "C" const char* sg_last_error(void) { return g_err.c_str(); }
extern "C" int sg_version(void) { return 100; }

#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) return fail(SG_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)
#define SGCALL(call)            \
  do {                          \
    int r_ = (call);            \
    if (r_ != SG_OK) return r_; \
  } while (0)
#define REQUIRE(cond, ...)                                  \
  do {                                                      \
    if (!(cond)) return fail(SG_ERR_INVALID, __VA_ARGS__);  \
  } while (0)

// ------------------------------------------------------------------------------------------------
// objects
// ------------------------------------------------------------------------------------------------
struct Box {
  int lo[2], hi[2];
  int nx() const { return hi[0] - lo[0] + 1; }
  int ny() const { return hi[1] - lo[1] + 1; }
  long long npts() const { return (long long)nx() * ny(); }
};
static inline int fdiv(int a, int r) { return a >= 0 ? a / r : -((-a + r - 1) / r); }

#define SG_MAXPART (1u << 20) // blocks of one residual sweep whose maxima are folded by k_max_partials (more: atomics)
struct sg_ctx {
  int device = 0, rank = 0, nranks = 1, num_sms = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  long long launches = 0;
  int relax_mode = 1;
  bool smem_attr_set = false; // dynamic shared-memory opt-in of the streaming kernels done for this context's device
  cudaStream_t comm_stream = nullptr;    // nranks > 1: halo exchanges that overlap the interior part of a sweep run here
  cudaEvent_t ev_comm[2] = {nullptr, nullptr};
  std::vector<struct sg_solver*> solvers; // live head solvers: their captured graphs (which hold NCCL kernels at N > 1) go before the communicator
  int tune[32] = {0}; // experiment knobs (sg_set_tuning): 0 rows per warp, 1 CTAs per SM of the fused sweep
  SgNccl nccl;
  // reduction scratch
  double* d_partial = nullptr;
  size_t partial_cap = 0;
  unsigned long long* d_maxpart = nullptr; // per-block maxima of the residual sweeps (k_max_partials), SG_MAXPART entries
  double* d_scalar = nullptr;            // [0..127] device scalars (as doubles / bit patterns)
  double* h_scalar = nullptr;            // pinned mirror
  double* h_stage = nullptr; size_t h_stage_cap = 0; // pinned staging for batched upload/download
  double* d_stage = nullptr; size_t d_stage_cap = 0;
  CopySeg* d_segs = nullptr; size_t segs_cap = 0;
  cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

// one peer of a copy plan that crosses ranks: segments packed into sendbuf / unpacked from recvbuf (ncclSend/ncclRecv)
struct PeerPlan {
  int rank = -1;
  CopySeg *d_send = nullptr, *d_recv = nullptr;
  int nsend = 0, nrecv = 0;
  size_t send_count = 0, recv_count = 0;
  double *sendbuf = nullptr, *recvbuf = nullptr;
};
struct sg_field;
struct sg_layout {
  sg_ctx* ctx;
  int nbox;
  std::vector<Box> boxes;
  std::vector<int> owner;
  Box domain;
  int periodic[2];
  bool has_local = false;
  Box patch;             // rank-local valid rectangle
  int nx = 0, ny = 0, pitch = 0, rows = 0;
  bool side_ghost[4] = {false, false, false, false}; // ghost data from a periodic image / neighbour rank
  bool side_domain[4] = {false, false, false, false}; // side lies on the problem-domain boundary
  int nbr[4] = {-1, -1, -1, -1};                      // neighbour rank for y sides (-1: none / self wrap)
  bool wrap_local[2] = {false, false};
  std::vector<sg_field*> ws; // lazily allocated work fields
  int refs = 1;
  // patch table: one entry (the merged rectangle) for uniform levels, one entry per box otherwise
  bool general = false;            // patches are the individual boxes (each with its own ghost ring)
  bool fast = false;               // one rectangle without coarse-fine sides: the streaming kernels of sg_kernels.cuh apply
  std::vector<PatchG> patches;
  std::vector<int> patch_of_box;   // -1: box owned by another rank; RECT: 0
  PatchG* d_patches = nullptr;
  size_t total = 0;                // doubles per component
  int max_nx = 0, max_ny = 0;
  bool any_phys = false;           // some patch side lies on a non-periodic domain boundary (else physical-BC kernels are skipped)
  struct Plan { CopySeg* d = nullptr; int n = 0; bool built = false; bool small = false; std::vector<PeerPlan> peers; };
  Plan ex_faces, ex_full[3];       // exchange plans (general path): face strips depth 1; all ghosts depth 1 / 2
  // spatial bins over the domain for "which patch holds cell (i,j)"
  int bin_s = 0, bin_nx = 0, bin_ny = 0;
  std::vector<std::vector<int>> bins;
  // the level as ALL ranks hold it: one entry per rank rectangle (merged levels) or per box (general levels), in an order
  // every rank computes identically -- copy plans between ranks are enumerated over it
  struct GPatch { Box b; int owner; int local; }; // local: index into `patches` when owner == this rank, else -1
  std::vector<GPatch> gp;
  int gbin_s = 0, gbin_nx = 0, gbin_ny = 0;
  std::vector<std::vector<int>> gbins;
  // fused per-patch smoother (k_gsrb_patch): ghost-cell records, built on first use.  0 not tried, 1 usable, -1 not usable
  // (a same-level neighbour box lives on another GPU, or a patch does not fit the shared-memory tile)
  int fused_state = 0;
  GRec* d_grec = nullptr;
  int* d_grec_start = nullptr;
  size_t fused_smem = 0;
  int fused_carveout = 0;
};
#define GEN_GX 2
#define GEN_GY 2

struct sg_field {
  sg_layout* lay;
  bool owns_layout = false; // createCoarser: the coarsened layout lives and dies with the field
  int ncomp, ng, cent;
  double* base = nullptr;
  size_t comp_stride = 0;
  double* p(int c = 0) const { return base + (size_t)c * comp_stride + (size_t)SG_YOFF * lay->pitch + SG_XOFF; } // fast path
  double* cb(int c = 0) const { return base + (size_t)c * comp_stride; } // component base for patch-table kernels
};

struct sg_factory {
  sg_ctx* ctx;
  int nlevels;
  std::vector<sg_layout*> grids;
  std::vector<int> ref_ratios;
  std::vector<std::array<double, 2>> dx;
  sg_bc bc;
  double alpha, beta;
  sg_params prm;
  std::vector<sg_field*> aCoef, bX, bY, B, Pi, zb, mask;
};

struct AmrLink;
struct FineLink;
struct sg_op {
  AmrLink* link = nullptr;   // coarse-fine interpolator etc. (levels above the base)
  FineLink* flink = nullptr; // flux register / covered cells with respect to the next finer level
  sg_ctx* ctx;
  sg_layout* lay;
  double dx[2];
  double alpha, beta;
  sg_bc bc;
  sg_params prm;
  sg_field *aCoef, *bX, *bY, *B, *Pi, *zb, *mask;
  bool owns_coefs = false;
  bool owns_layout = false;
  int level = 0, depth = 0;
  bool update_operator = false;
  bool mask_needed = true; // false: no negative ice-mask entry on this level (set by op_scan_mask)
};

struct sg_solver {
  sg_ctx* ctx;
  sg_factory* fac;
  int num_levels;
  std::vector<sg_op*> ops;                          // MG depths of level 0 (ops[0] is the base AMR level's operator)
  std::vector<sg_field*> phi, rhs, save, tmp;       // per depth (depth 0 entries unused except tmp/resid)
  sg_field* resid = nullptr;                        // = aresid[0]
  // AMR levels (index = level; entry 0 of aops aliases ops[0])
  std::vector<sg_op*> aops;
  std::vector<sg_field*> aresid, acorr, atmp, ascratch, aresC;
  std::vector<sg_field*> asave;                     // level l > 0: the coarser level's phi under and around level l before the coarse solve (coarsened-fine layout)
  // one FAS V-cycle + residual norm captured as a CUDA graph (launch-bound levels: ~110 launches become one)
  bool coefs_averaged = false;                      // this V-cycle's depth >= 1 face coefficients are already averaged (average_all_depths)
  cudaGraphExec_t gexec = nullptr;
  std::vector<long long> gkey, warm_key;
  long long glaunches = 0;
};

// ------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------
#define LAUNCH(ctx, kernel, grid, block, ...)                    \
  do {                                                           \
    kernel<<<(grid), (block), 0, (ctx)->stream>>>(__VA_ARGS__);  \
    (ctx)->launches++;                                           \
  } while (0)
static inline dim3 grid2(int nx, int ny, dim3 b) { return dim3((nx + b.x - 1) / b.x, (ny + b.y - 1) / b.y); }
static const dim3 B2D(32, 8);

static PhysP phys(const sg_params& p) {
  PhysP q;
  q.A = p.A; q.cutOffbr = p.cutOffbr; q.maxOffbr = p.maxOffbr; q.omega = p.omega; q.nu = p.nu;
  q.cutOffBcoef = p.cutOffBcoef; q.use_NL = p.use_NL; q.use_mask_grad = p.use_mask_grad;
  return q;
}

static Geom make_geom(const sg_layout* L, const sg_bc* bc) {
  Geom g;
  g.nx = L->nx; g.ny = L->ny; g.pitch = L->pitch;
  g.glo0 = L->patch.lo[0]; g.glo1 = L->patch.lo[1];
  g.dlo0 = L->domain.lo[0]; g.dlo1 = L->domain.lo[1]; g.dhi0 = L->domain.hi[0]; g.dhi1 = L->domain.hi[1];
  for (int s = 0; s < 4; s++) {
    int dir = s >> 1, side = s & 1;
    g.bcval[s] = 0.0;
    if (L->side_ghost[s]) g.kind[s] = SK_GHOST;
    else if (L->side_domain[s]) {
      int type = bc ? (side ? bc->hi_type[dir] : bc->lo_type[dir]) : -1;
      g.kind[s] = type == 0 ? SK_PHYS_DIRI : type == 1 ? SK_PHYS_NEUM : SK_PHYS_NONE;
      if (bc) g.bcval[s] = side ? bc->hi_val[dir] : bc->lo_val[dir];
    } else g.kind[s] = SK_FROZEN;
  }
  return g;
}

static OpArgs make_args(const sg_op* op) {
  OpArgs a;
  a.g = make_geom(op->lay, &op->bc);
  a.prm = phys(op->prm);
  a.alpha = op->alpha; a.beta = op->beta;
  a.dx0 = op->dx[0]; a.dx1 = op->dx[1];
  a.dxi0 = 1.0 / (op->dx[0] * op->dx[0]);
  a.dxi1 = 1.0 / (op->dx[1] * op->dx[1]);
  a.has_a = op->alpha != 0.0;
  a.use_mask = op->mask_needed ? 1 : 0;
  a.aC = op->aCoef ? op->aCoef->p() : nullptr;
  a.bX = op->bX->p(); a.bY = op->bY->p();
  a.B = op->B->p(); a.Pi = op->Pi->p(); a.zb = op->zb->p(); a.mask = op->mask->p();
  return a;
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
extern "C" int sg_nccl_unique_id(void* out128) {
  REQUIRE(out128, "sg_nccl_unique_id: null");
  return sgnccl_unique_id(out128, g_err);
}

extern "C" int sg_ctx_create(sg_ctx** out, int device, int rank, int nranks, const void* nccl_unique_id) {
  REQUIRE(out, "sg_ctx_create: null out");
  REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "sg_ctx_create: bad rank %d/%d", rank, nranks);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(SG_ERR_CUDA, "sg_ctx_create: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
  REQUIRE(device >= 0 && device < ndev, "sg_ctx_create: device %d of %d", device, ndev);
  CK(cudaSetDevice(device));
  sg_ctx* c = new sg_ctx();
  c->device = device; c->rank = rank; c->nranks = nranks;
  CK(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device));
  CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CK(cudaMalloc(&c->d_scalar, 128 * sizeof(double)));
  CK(cudaMalloc(&c->d_maxpart, SG_MAXPART * sizeof(unsigned long long)));
  CK(cudaMemsetAsync(c->d_scalar, 0, 128 * sizeof(double), c->stream));
  CK(cudaMallocHost(&c->h_scalar, 128 * sizeof(double)));
  if (nranks > 1) {
    REQUIRE(nccl_unique_id, "sg_ctx_create: nranks > 1 needs an NCCL unique id");
    int r = c->nccl.init(nccl_unique_id, rank, nranks, g_err);
    if (r != SG_OK) return r;
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi));
    CK(cudaEventCreateWithFlags(&c->ev_comm[0], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_comm[1], cudaEventDisableTiming));
  }
  *out = c;
  return SG_OK;
}
struct sg_layout;
static void gap_cache_forget(const sg_ctx* ctx, const sg_layout* L);
static void drop_solver_graphs(sg_ctx* c);
extern "C" int sg_ctx_destroy(sg_ctx* c) {
  if (!c) return SG_OK;
  cudaSetDevice(c->device);
  gap_cache_forget(c, nullptr); // cached implicit gap-height solvers own device fields on this context
  cudaStreamSynchronize(c->stream);
  drop_solver_graphs(c); // ncclCommDestroy waits for every graph that captured one of its kernels
  if (c->comm_stream) { cudaStreamSynchronize(c->comm_stream); cudaStreamDestroy(c->comm_stream); }
  for (int k = 0; k < 2; k++) if (c->ev_comm[k]) cudaEventDestroy(c->ev_comm[k]);
  c->nccl.destroy();
  cudaFree(c->d_partial); cudaFree(c->d_maxpart); cudaFree(c->d_scalar); cudaFreeHost(c->h_scalar);
  cudaFreeHost(c->h_stage); cudaFree(c->d_stage); cudaFree(c->d_segs);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
  return SG_OK;
}
extern "C" int sg_ctx_sync(sg_ctx* c) {
  REQUIRE(c, "null ctx");
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaGetLastError());
  return SG_OK;
}
extern "C" int sg_ctx_set_stream(sg_ctx* c, void* s) {
  REQUIRE(c, "null ctx");
  CK(cudaStreamSynchronize(c->stream));
  if (c->own_stream) cudaStreamDestroy(c->stream);
  c->stream = (cudaStream_t)s;
  c->own_stream = false;
  return SG_OK;
}
extern "C" int sg_ctx_kernel_launches(sg_ctx* c, long long* out) {
  REQUIRE(c && out, "null");
  *out = c->launches;
  return SG_OK;
}
extern "C" int sg_ctx_event_record(sg_ctx* c, int slot) {
  REQUIRE(c && slot >= 0 && slot < 8, "sg_ctx_event_record: slot 0..7");
  if (!c->ev[slot]) CK(cudaEventCreate(&c->ev[slot]));
  CK(cudaEventRecord(c->ev[slot], c->stream));
  return SG_OK;
}
extern "C" int sg_ctx_event_elapsed_ms(sg_ctx* c, int slot0, int slot1, double* ms) {
  REQUIRE(c && ms && slot0 >= 0 && slot0 < 8 && slot1 >= 0 && slot1 < 8 && c->ev[slot0] && c->ev[slot1], "sg_ctx_event_elapsed_ms: bad slots");
  CK(cudaEventSynchronize(c->ev[slot1]));
  float f = 0;
  CK(cudaEventElapsedTime(&f, c->ev[slot0], c->ev[slot1]));
  *ms = f;
  return SG_OK;
}
extern "C" int sg_set_tuning(sg_ctx* c, int key, int value) {
  REQUIRE(c && key >= 0 && key < 32, "sg_set_tuning: key 0..31");
  c->tune[key] = value;
  return SG_OK;
}
extern "C" int sg_set_relax_mode(sg_ctx* c, int mode) {
  REQUIRE(c && mode >= 0 && mode <= 5, "sg_set_relax_mode");
  c->relax_mode = mode;
  return SG_OK;
}

// ------------------------------------------------------------------------------------------------
// layouts
// ------------------------------------------------------------------------------------------------
static inline bool in_domain_p(const sg_layout* L, int i, int j) {
  if (!L->periodic[0] && (i < L->domain.lo[0] || i > L->domain.hi[0])) return false;
  if (!L->periodic[1] && (j < L->domain.lo[1] || j > L->domain.hi[1])) return false;
  return true;
}
static inline Box patch_box(const PatchG& g) {
  Box b;
  b.lo[0] = g.glo0; b.lo[1] = g.glo1; b.hi[0] = g.glo0 + g.nx - 1; b.hi[1] = g.glo1 + g.ny - 1;
  return b;
}
// element offset (from the component base) of global cell (gi,gj) inside patch g's array (ghost region included)
static inline long long patch_off(const PatchG& g, int gi, int gj) { return g.off + (long long)(gj - g.glo1) * g.pitch + (gi - g.glo0); }

static int layout_tables(sg_layout* L) {
  sg_ctx* c = L->ctx;
  if (!L->has_local) return SG_OK;
  L->max_nx = L->max_ny = 0;
  for (const PatchG& g : L->patches) {
    L->max_nx = std::max(L->max_nx, g.nx); L->max_ny = std::max(L->max_ny, g.ny);
    L->any_phys = L->any_phys || g.phys[0] || g.phys[1] || g.phys[2] || g.phys[3];
  }
  CK(cudaMalloc(&L->d_patches, L->patches.size() * sizeof(PatchG)));
  CK(cudaMemcpyAsync(L->d_patches, L->patches.data(), L->patches.size() * sizeof(PatchG), cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  L->bin_s = std::max(16, std::max(L->max_nx, L->max_ny));
  L->bin_nx = (L->domain.hi[0] - L->domain.lo[0]) / L->bin_s + 1;
  L->bin_ny = (L->domain.hi[1] - L->domain.lo[1]) / L->bin_s + 1;
  L->bins.assign((size_t)L->bin_nx * L->bin_ny, std::vector<int>());
  for (int k = 0; k < (int)L->patches.size(); k++) {
    const PatchG& g = L->patches[k];
    int bx0 = (g.glo0 - L->domain.lo[0]) / L->bin_s, bx1 = (g.glo0 + g.nx - 1 - L->domain.lo[0]) / L->bin_s;
    int by0 = (g.glo1 - L->domain.lo[1]) / L->bin_s, by1 = (g.glo1 + g.ny - 1 - L->domain.lo[1]) / L->bin_s;
    for (int by = by0; by <= by1; by++)
      for (int bx = bx0; bx <= bx1; bx++) L->bins[(size_t)by * L->bin_nx + bx].push_back(k);
  }
  return SG_OK;
}

// host part of a layout: ownership, merged patch / per-box patches, neighbour ranks (no device calls)
static int layout_host(sg_layout* L) {
  sg_ctx* c = L->ctx;
  // per-rank bounding rectangles; a rank whose boxes tile its rectangle exactly stores the level as ONE merged patch
  std::vector<Box> rp(c->nranks);
  std::vector<long long> area(c->nranks, 0);
  std::vector<bool> has(c->nranks, false);
  for (int b = 0; b < L->nbox; b++) {
    int r = L->owner[b];
    REQUIRE(r >= 0 && r < c->nranks, "layout: box %d owned by rank %d of %d", b, r, c->nranks);
    const Box& bx = L->boxes[b];
    REQUIRE(bx.nx() > 0 && bx.ny() > 0, "layout: empty box %d", b);
    if (!has[r]) { rp[r] = bx; has[r] = true; }
    else
      for (int d = 0; d < 2; d++) { rp[r].lo[d] = std::min(rp[r].lo[d], bx.lo[d]); rp[r].hi[d] = std::max(rp[r].hi[d], bx.hi[d]); }
    area[r] += bx.npts();
  }
  bool tiles = true;
  long long covered = 0;
  for (int r = 0; r < c->nranks; r++) {
    if (has[r] && area[r] != rp[r].npts()) tiles = false;
    covered += area[r];
    // across ranks only full-width y-strips of a level that covers the whole domain are stored merged (NCCL row exchange)
    if (c->nranks > 1 && has[r] && (rp[r].lo[0] != L->domain.lo[0] || rp[r].hi[0] != L->domain.hi[0])) tiles = false;
  }
  // A level is stored merged only when its boxes cover the whole domain (every base level does).  A refined level whose boxes
  // happen to tile a rectangle stays a list of per-box patches: merged storage has no private ghost cells between boxes, and
  // AMRProlongS_2 reads such cells where two boxes meet on a physical boundary (the corner ghost cell of the coarsened scratch outside
  // the domain, which neither the boundary condition nor the corner exchange fills, src/AMRNonLinearPoissonOp.cpp:1155-1172).
  if (covered != L->domain.npts()) tiles = false;
  L->has_local = has[c->rank];
  L->patch_of_box.assign(L->nbox, -1);
  L->gp.clear();
  if (!tiles) {
    // refined AMR level: every box is its own patch with a private ghost ring (Chombo's own data model); a rank stores the
    // boxes it owns, ghost filling and inter-level copies run over copy plans that may cross ranks
    L->general = true; L->fast = false;
    L->patch = has[c->rank] ? rp[c->rank] : L->boxes[0];
    L->nx = L->patch.nx(); L->ny = L->patch.ny(); L->pitch = 0; L->rows = 0;
    size_t tot = 0;
    for (int b = 0; b < L->nbox; b++) {
      const Box& bx = L->boxes[b];
      sg_layout::GPatch q;
      q.b = bx; q.owner = L->owner[b]; q.local = -1;
      if (L->owner[b] != c->rank) { L->gp.push_back(q); continue; }
      q.local = (int)L->patches.size();
      L->gp.push_back(q);
      PatchG g;
      g.nx = bx.nx(); g.ny = bx.ny(); g.glo0 = bx.lo[0]; g.glo1 = bx.lo[1];
      g.pitch = ((g.nx + 2 * GEN_GX + 2 + 1) / 2) * 2;
      int rows = g.ny + 2 * GEN_GY + 2;
      g.off = (long long)tot + (long long)GEN_GY * g.pitch + GEN_GX;
      for (int s = 0; s < 4; s++) {
        int dir = s >> 1, side = s & 1;
        bool ondom = side ? bx.hi[dir] == L->domain.hi[dir] : bx.lo[dir] == L->domain.lo[dir];
        g.phys[s] = ondom && !L->periodic[dir];
      }
      tot += (size_t)rows * g.pitch;
      L->patch_of_box[b] = (int)L->patches.size();
      L->patches.push_back(g);
    }
    L->total = tot;
    return SG_OK;
  }
  for (int r = 0; r < c->nranks; r++)
    if (has[r]) {
      sg_layout::GPatch q;
      q.b = rp[r]; q.owner = r; q.local = r == c->rank ? 0 : -1;
      L->gp.push_back(q);
    }
  if (!L->has_local) return SG_OK;
  L->patch = rp[c->rank];
  L->nx = L->patch.nx(); L->ny = L->patch.ny();
  L->pitch = ((L->nx + 2 * SG_XOFF + 15) / 16) * 16;
  L->rows = L->ny + SG_YOFF + SG_YTOP;
  for (int s = 0; s < 4; s++) {
    int dir = s >> 1, side = s & 1;
    bool ondom = side ? L->patch.hi[dir] == L->domain.hi[dir] : L->patch.lo[dir] == L->domain.lo[dir];
    L->side_domain[s] = ondom;
    L->side_ghost[s] = false;
    L->nbr[s] = -1;
    bool spans = L->patch.lo[dir] == L->domain.lo[dir] && L->patch.hi[dir] == L->domain.hi[dir];
    if (ondom && L->periodic[dir]) {
      L->side_ghost[s] = true;
      if (spans) L->wrap_local[dir] = true;
    }
    if (!(ondom && L->periodic[dir] && spans) && (!ondom || L->periodic[dir])) {
      // need a neighbouring rank's patch across this side (only full-width strips in y are supported)
      if (dir == 0) {
        if (!ondom) {
          // level does not reach the domain side and nobody is next to it: coarse-fine boundary
          bool found = false;
          for (int r = 0; r < c->nranks; r++)
            if (r != c->rank && has[r]) {
              bool adj = side ? rp[r].lo[0] == L->patch.hi[0] + 1 : rp[r].hi[0] == L->patch.lo[0] - 1;
              if (adj && rp[r].lo[1] <= L->patch.hi[1] && rp[r].hi[1] >= L->patch.lo[1]) found = true;
            }
          if (found) return fail(SG_ERR_UNSUPPORTED, "layout: box partition splits the x direction across ranks");
          L->side_ghost[s] = false;
        } else return fail(SG_ERR_UNSUPPORTED, "layout: periodic x across ranks");
      } else {
        int want = side ? L->patch.hi[1] + 1 : L->patch.lo[1] - 1;
        int ny_dom = L->domain.hi[1] - L->domain.lo[1] + 1;
        if (ondom) want = side ? want - ny_dom : want + ny_dom;
        int found = -1;
        for (int r = 0; r < c->nranks; r++)
          if (has[r] && r != c->rank) {
            bool adj = side ? rp[r].lo[1] == want : rp[r].hi[1] == want;
            if (adj && rp[r].lo[0] == L->patch.lo[0] && rp[r].hi[0] == L->patch.hi[0]) found = r;
            else if (adj && rp[r].lo[0] <= L->patch.hi[0] && rp[r].hi[0] >= L->patch.lo[0])
              return fail(SG_ERR_UNSUPPORTED, "layout: neighbouring rank patches must share the x extent");
          }
        if (found >= 0) { L->nbr[s] = found; L->side_ghost[s] = true; }
        else REQUIRE(!ondom, "layout: periodic y side without an owner");
      }
    }
  }
  {
    PatchG g;
    g.off = (long long)SG_YOFF * L->pitch + SG_XOFF;
    g.pitch = L->pitch; g.nx = L->nx; g.ny = L->ny; g.glo0 = L->patch.lo[0]; g.glo1 = L->patch.lo[1];
    L->fast = true;
    for (int s = 0; s < 4; s++) {
      g.phys[s] = L->side_domain[s] && !L->periodic[s >> 1];
      if (!L->side_domain[s] && !L->side_ghost[s]) L->fast = false; // coarse-fine side: ghost cells hold interpolated data
    }
    L->patches.push_back(g);
    L->total = (size_t)L->rows * L->pitch;
    for (int b = 0; b < L->nbox; b++)
      if (L->owner[b] == c->rank) L->patch_of_box[b] = 0;
  }
  return SG_OK;
}
static void layout_global_bins(sg_layout* L) {
  int smax = 16;
  for (const auto& q : L->gp) smax = std::max(smax, std::max(q.b.nx(), q.b.ny()));
  L->gbin_s = smax;
  L->gbin_nx = (L->domain.hi[0] - L->domain.lo[0]) / smax + 1;
  L->gbin_ny = (L->domain.hi[1] - L->domain.lo[1]) / smax + 1;
  L->gbins.assign((size_t)L->gbin_nx * L->gbin_ny, std::vector<int>());
  for (int k = 0; k < (int)L->gp.size(); k++) {
    const Box& b = L->gp[k].b;
    int bx0 = (b.lo[0] - L->domain.lo[0]) / smax, bx1 = (b.hi[0] - L->domain.lo[0]) / smax;
    int by0 = (b.lo[1] - L->domain.lo[1]) / smax, by1 = (b.hi[1] - L->domain.lo[1]) / smax;
    for (int by = by0; by <= by1; by++)
      for (int bx = bx0; bx <= bx1; bx++) L->gbins[(size_t)by * L->gbin_nx + bx].push_back(k);
  }
}
static int layout_finish(sg_layout* L) {
  SGCALL(layout_host(L));
  layout_global_bins(L);
  return layout_tables(L);
}

// ---- host-only views of the partition (no CUDA device needed): what the gloo tests and a driver's planning use ----
// LoadBalance for the box-wise strip partition (absent Chombo LoadBalance as called at src/AmrHydro.cpp:4847,4929):
// contiguous runs of the (y-major sorted) box list with equal cell counts, cut only between whole rows of boxes so that
// every rank's boxes tile one rectangle.
extern "C" int sg_partition_boxes(int nbox, const int* boxes, int nranks, int* owner_out) {
  REQUIRE(nbox > 0 && boxes && nranks >= 1 && owner_out, "sg_partition_boxes: bad arguments");
  // distinct box rows (by lo1), cells per row
  std::map<int, long long> row_cells;
  long long total = 0;
  for (int b = 0; b < nbox; b++) {
    long long n = (long long)(boxes[4 * b + 2] - boxes[4 * b] + 1) * (boxes[4 * b + 3] - boxes[4 * b + 1] + 1);
    row_cells[boxes[4 * b + 1]] += n;
    total += n;
  }
  REQUIRE((int)row_cells.size() >= nranks, "sg_partition_boxes: %d rows of boxes cannot be cut into %d strips", (int)row_cells.size(), nranks);
  std::map<int, int> row_owner;
  long long acc = 0;
  int r = 0, rows_left = (int)row_cells.size();
  for (auto& kv : row_cells) {
    // move on when this rank has its share, but keep at least one row for every remaining rank
    int ranks_left = nranks - r;
    if (r < nranks - 1 && (acc >= total * (r + 1) / nranks || rows_left < ranks_left)) r++;
    row_owner[kv.first] = r;
    acc += kv.second;
    rows_left--;
  }
  for (int b = 0; b < nbox; b++) owner_out[b] = row_owner[boxes[4 * b + 1]];
  return SG_OK;
}
// the rank-local rectangle, the neighbour ranks across its four sides (-1: none; a periodic image that wraps onto the same
// rank is "none" here, it needs no message) and the doubles per halo row a field exchange moves
extern "C" int sg_partition_describe(int nbox, const int* boxes, const int* owner, const int domain[4], const int periodic[2], int rank,
                                     int nranks, int patch_out[4], int nbr_out[4], long long* halo_row_doubles) {
  REQUIRE(nbox > 0 && boxes && domain && periodic && patch_out && nbr_out && rank >= 0 && rank < nranks, "sg_partition_describe: bad arguments");
  sg_ctx fake;
  fake.rank = rank; fake.nranks = nranks;
  sg_layout L;
  L.ctx = &fake; L.nbox = nbox;
  L.boxes.resize(nbox); L.owner.resize(nbox);
  for (int b = 0; b < nbox; b++) {
    L.boxes[b].lo[0] = boxes[4 * b]; L.boxes[b].lo[1] = boxes[4 * b + 1]; L.boxes[b].hi[0] = boxes[4 * b + 2]; L.boxes[b].hi[1] = boxes[4 * b + 3];
    L.owner[b] = owner ? owner[b] : 0;
  }
  L.domain.lo[0] = domain[0]; L.domain.lo[1] = domain[1]; L.domain.hi[0] = domain[2]; L.domain.hi[1] = domain[3];
  L.periodic[0] = periodic[0]; L.periodic[1] = periodic[1];
  SGCALL(layout_host(&L));
  REQUIRE(L.has_local, "sg_partition_describe: rank %d owns no box", rank);
  patch_out[0] = L.patch.lo[0]; patch_out[1] = L.patch.lo[1]; patch_out[2] = L.patch.hi[0]; patch_out[3] = L.patch.hi[1];
  for (int s2 = 0; s2 < 4; s2++) nbr_out[s2] = L.nbr[s2];
  if (halo_row_doubles) *halo_row_doubles = L.pitch;
  return SG_OK;
}

extern "C" int sg_layout_create(sg_ctx* ctx, sg_layout** out, int nbox, const int* boxes, const int* owner,
                                const int domain[4], const int periodic[2]) {
  REQUIRE(ctx && out && boxes && domain && periodic && nbox > 0, "sg_layout_create: bad arguments");
  sg_layout* L = new sg_layout();
  L->ctx = ctx; L->nbox = nbox;
  L->boxes.resize(nbox); L->owner.resize(nbox);
  for (int b = 0; b < nbox; b++) {
    L->boxes[b].lo[0] = boxes[4 * b]; L->boxes[b].lo[1] = boxes[4 * b + 1];
    L->boxes[b].hi[0] = boxes[4 * b + 2]; L->boxes[b].hi[1] = boxes[4 * b + 3];
    L->owner[b] = owner ? owner[b] : 0;
  }
  L->domain.lo[0] = domain[0]; L->domain.lo[1] = domain[1]; L->domain.hi[0] = domain[2]; L->domain.hi[1] = domain[3];
  L->periodic[0] = periodic[0]; L->periodic[1] = periodic[1];
  int r = layout_finish(L);
  if (r != SG_OK) { delete L; return r; }
  *out = L;
  return SG_OK;
}
extern "C" int sg_layout_coarsenable(const sg_layout* L, int ratio, int* out) {
  REQUIRE(L && out && ratio >= 1, "sg_layout_coarsenable");
  *out = 1;
  for (const Box& b : L->boxes)
    for (int d = 0; d < 2; d++)
      if (fdiv(b.lo[d], ratio) * ratio != b.lo[d] || fdiv(b.hi[d], ratio) * ratio + ratio - 1 != b.hi[d]) { *out = 0; return SG_OK; }
  return SG_OK;
}
extern "C" int sg_layout_coarsen(sg_layout* L, int ratio, sg_layout** out) {
  REQUIRE(L && out && ratio >= 1, "sg_layout_coarsen");
  sg_layout* C = new sg_layout();
  C->ctx = L->ctx; C->nbox = L->nbox;
  C->boxes.resize(L->nbox); C->owner = L->owner;
  for (int b = 0; b < L->nbox; b++)
    for (int d = 0; d < 2; d++) { C->boxes[b].lo[d] = fdiv(L->boxes[b].lo[d], ratio); C->boxes[b].hi[d] = fdiv(L->boxes[b].hi[d], ratio); }
  for (int d = 0; d < 2; d++) { C->domain.lo[d] = fdiv(L->domain.lo[d], ratio); C->domain.hi[d] = fdiv(L->domain.hi[d], ratio); }
  C->periodic[0] = L->periodic[0]; C->periodic[1] = L->periodic[1];
  int r = layout_finish(C);
  if (r != SG_OK) { delete C; return r; }
  *out = C;
  return SG_OK;
}
extern "C" int sg_layout_nbox(const sg_layout* L, int* nbox) {
  REQUIRE(L && nbox, "sg_layout_nbox");
  *nbox = L->nbox;
  return SG_OK;
}
extern "C" int sg_field_destroy(sg_field* f);
static void copy_plans_forget(const sg_layout* L);
static void gap_cache_forget(const sg_ctx* ctx, const sg_layout* L);
static void plan_free(sg_layout::Plan& P);
extern "C" int sg_layout_destroy(sg_layout* L) {
  if (!L) return SG_OK;
  gap_cache_forget(nullptr, L);
  for (sg_field* w : L->ws) sg_field_destroy(w);
  copy_plans_forget(L);
  cudaFree(L->d_patches);
  cudaFree(L->d_grec); cudaFree(L->d_grec_start);
  plan_free(L->ex_faces);
  for (int k = 0; k < 3; k++) plan_free(L->ex_full[k]);
  delete L;
  return SG_OK;
}

// ------------------------------------------------------------------------------------------------
// fields
// ------------------------------------------------------------------------------------------------
extern "C" int sg_field_create(sg_layout* L, sg_field** out, int ncomp, int nghost, int centering) {
  REQUIRE(L && out && ncomp >= 1 && nghost >= 0 && nghost <= 2 && centering >= 0 && centering <= 2, "sg_field_create: bad arguments");
  sg_field* f = new sg_field();
  f->lay = L; f->ncomp = ncomp; f->ng = nghost; f->cent = centering;
  if (L->has_local) {
    f->comp_stride = L->total;
    CK(cudaMalloc(&f->base, f->comp_stride * ncomp * sizeof(double)));
    CK(cudaMemsetAsync(f->base, 0, f->comp_stride * ncomp * sizeof(double), L->ctx->stream));
  }
  *out = f;
  return SG_OK;
}
extern "C" int sg_field_destroy(sg_field* f) {
  if (!f) return SG_OK;
  if (f->base) {
    cudaStreamSynchronize(f->lay->ctx->stream);
    cudaFree(f->base);
  }
  if (f->owns_layout) sg_layout_destroy(f->lay);
  delete f;
  return SG_OK;
}
// work field k of a layout (same shape as every other field of the layout)
static int ws_field(sg_layout* L, int k, int ncomp, sg_field** out) {
  if ((int)L->ws.size() <= k) L->ws.resize(k + 1, nullptr);
  if (!L->ws[k] || L->ws[k]->ncomp < ncomp) {
    if (L->ws[k]) sg_field_destroy(L->ws[k]);
    SGCALL(sg_field_create(L, &L->ws[k], ncomp, 2, SG_CELL));
  }
  *out = L->ws[k];
  return SG_OK;
}

// rectangle of box b's FArrayBox in the field's centering, ghosts included
static Box fab_rect(const sg_field* f, int b, int ng) {
  Box r = f->lay->boxes[b];
  for (int d = 0; d < 2; d++) { r.lo[d] -= ng; r.hi[d] += ng; }
  if (f->cent == SG_XFACE) r.hi[0] += 1;
  if (f->cent == SG_YFACE) r.hi[1] += 1;
  return r;
}
static Box patch_rect(const sg_field* f) {
  Box r = f->lay->patch;
  if (f->cent == SG_XFACE) r.hi[0] += 1;
  if (f->cent == SG_YFACE) r.hi[1] += 1;
  return r;
}
// pieces of box b's FAB that are copied on upload: the valid region plus the ghost cells lying outside the patch
static int upload_rects(const sg_field* f, int b, Box out[5]) {
  int n = 0;
  if (f->lay->general) { out[0] = fab_rect(f, b, f->ng); return 1; } // every box owns its ghost ring
  Box F = fab_rect(f, b, f->ng), V = fab_rect(f, b, 0), Pv = patch_rect(f);
  out[n++] = V;
  if (f->ng == 0) return n;
  // Ghost cells outside the merged patch exist once, but several boxes' FArrayBoxes overlap there: a cell of the strip
  // below/above (left/right of) the patch is taken from the box whose valid columns (rows) contain it, and a corner of
  // the patch from the box at that corner -- the box for which the cell is a FACE ghost, not a corner ghost.
  Box r;
  const int x0 = F.lo[0] < Pv.lo[0] ? F.lo[0] : V.lo[0], x1 = F.hi[0] > Pv.hi[0] ? F.hi[0] : V.hi[0];
  if (F.lo[1] < Pv.lo[1]) { r = F; r.lo[0] = x0; r.hi[0] = x1; r.hi[1] = Pv.lo[1] - 1; out[n++] = r; }
  if (F.hi[1] > Pv.hi[1]) { r = F; r.lo[0] = x0; r.hi[0] = x1; r.lo[1] = Pv.hi[1] + 1; out[n++] = r; }
  if (F.lo[0] < Pv.lo[0]) { r = V; r.lo[0] = F.lo[0]; r.hi[0] = Pv.lo[0] - 1; out[n++] = r; }
  if (F.hi[0] > Pv.hi[0]) { r = V; r.hi[0] = F.hi[0]; r.lo[0] = Pv.hi[0] + 1; out[n++] = r; }
  return n;
}
// element offset from the component base of global index (gi,gj) in the array of the patch that holds box b
static inline ptrdiff_t dev_off(const sg_field* f, int b, int gi, int gj) {
  return (ptrdiff_t)patch_off(f->lay->patches[f->lay->patch_of_box[b]], gi, gj);
}
static inline int box_pitch(const sg_field* f, int b) { return f->lay->patches[f->lay->patch_of_box[b]].pitch; }

extern "C" int sg_field_upload_box(sg_field* f, int box, const double* host) {
  REQUIRE(f && host && box >= 0 && box < f->lay->nbox, "sg_field_upload_box: bad arguments");
  sg_layout* L = f->lay;
  REQUIRE(L->owner[box] == L->ctx->rank, "sg_field_upload_box: box %d is owned by rank %d", box, L->owner[box]);
  Box F = fab_rect(f, box, f->ng);
  Box rs[5];
  int n = upload_rects(f, box, rs);
  size_t fabn = (size_t)F.nx() * F.ny();
  for (int c = 0; c < f->ncomp; c++)
    for (int k = 0; k < n; k++) {
      const Box& r = rs[k];
      const double* src = host + c * fabn + (size_t)(r.lo[1] - F.lo[1]) * F.nx() + (r.lo[0] - F.lo[0]);
      double* dst = f->cb(c) + dev_off(f, box, r.lo[0], r.lo[1]);
      CK(cudaMemcpy2DAsync(dst, (size_t)box_pitch(f, box) * 8, src, (size_t)F.nx() * 8, (size_t)r.nx() * 8, r.ny(), cudaMemcpyHostToDevice, L->ctx->stream));
    }
  CK(cudaStreamSynchronize(L->ctx->stream)); // host buffer may be reused by the caller
  return SG_OK;
}
extern "C" int sg_field_download_box(const sg_field* f, int box, double* host) {
  REQUIRE(f && host && box >= 0 && box < f->lay->nbox, "sg_field_download_box: bad arguments");
  sg_layout* L = f->lay;
  REQUIRE(L->owner[box] == L->ctx->rank, "sg_field_download_box: box %d is owned by rank %d", box, L->owner[box]);
  Box F = fab_rect(f, box, f->ng);
  size_t fabn = (size_t)F.nx() * F.ny();
  for (int c = 0; c < f->ncomp; c++) {
    const double* src = f->cb(c) + dev_off(f, box, F.lo[0], F.lo[1]);
    CK(cudaMemcpy2DAsync(host + c * fabn, (size_t)F.nx() * 8, src, (size_t)box_pitch(f, box) * 8, (size_t)F.nx() * 8, F.ny(), cudaMemcpyDeviceToHost, L->ctx->stream));
  }
  CK(cudaStreamSynchronize(L->ctx->stream));
  return SG_OK;
}

// segment tables for moving all owned FABs (packed consecutively in box order) to/from the patch array
static size_t build_segs(const sg_field* f, bool upload, std::vector<CopySeg>& segs, std::vector<size_t>& offs) {
  sg_layout* L = f->lay;
  sg_ctx* c = L->ctx;
  size_t total = 0;
  offs.assign(L->nbox, 0);
  for (int b = 0; b < L->nbox; b++) {
    if (L->owner[b] != c->rank) continue;
    Box F = fab_rect(f, b, f->ng);
    offs[b] = total;
    Box rs[5];
    int n = 1;
    if (upload) n = upload_rects(f, b, rs); else rs[0] = F;
    for (int cc = 0; cc < f->ncomp; cc++)
      for (int k = 0; k < n; k++) {
        CopySeg s;
        long long packed = (long long)(total + (size_t)cc * F.nx() * F.ny() + (size_t)(rs[k].lo[1] - F.lo[1]) * F.nx() + (rs[k].lo[0] - F.lo[0]));
        long long dev = (long long)((f->cb(cc) - f->base) + dev_off(f, b, rs[k].lo[0], rs[k].lo[1]));
        s.so = upload ? packed : dev; s.dofs = upload ? dev : packed;
        s.nx = rs[k].nx(); s.ny = rs[k].ny();
        s.sp = upload ? F.nx() : box_pitch(f, b); s.dp = upload ? box_pitch(f, b) : F.nx();
        segs.push_back(s);
      }
    total += (size_t)F.nx() * F.ny() * f->ncomp;
  }
  return total;
}
static int scatter_from_stage(sg_field* f, const std::vector<CopySeg>& segs, const double* host_packed, size_t total) {
  sg_ctx* c = f->lay->ctx;
  CK(cudaMemcpyAsync(c->d_stage, host_packed, total * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->d_segs, segs.data(), segs.size() * sizeof(CopySeg), cudaMemcpyHostToDevice, c->stream));
  LAUNCH(c, k_copy_segs, dim3((unsigned)segs.size(), 4), 256, f->base, c->d_stage, c->d_segs, (int)segs.size());
  CK(cudaStreamSynchronize(c->stream)); // segs / caller's buffer may go away
  return SG_OK;
}
static int ensure_dev_stage(sg_ctx* c, size_t ndoubles, size_t nsegs, bool host_too) {
  if (c->d_stage_cap < ndoubles) {
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(c->d_stage); c->d_stage = nullptr;
    CK(cudaMalloc(&c->d_stage, ndoubles * sizeof(double)));
    c->d_stage_cap = ndoubles;
  }
  if (host_too && c->h_stage_cap < ndoubles) {
    CK(cudaStreamSynchronize(c->stream));
    cudaFreeHost(c->h_stage); c->h_stage = nullptr;
    CK(cudaMallocHost(&c->h_stage, ndoubles * sizeof(double)));
    c->h_stage_cap = ndoubles;
  }
  if (c->segs_cap < nsegs) {
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(c->d_segs);
    CK(cudaMalloc(&c->d_segs, nsegs * sizeof(CopySeg)));
    c->segs_cap = nsegs;
  }
  return SG_OK;
}

// batched upload: pack every owned FAB into pinned memory, one H2D copy, one scatter kernel
extern "C" int sg_field_upload(sg_field* f, const double* const* fabs) {
  REQUIRE(f && fabs, "sg_field_upload: bad arguments");
  sg_layout* L = f->lay;
  sg_ctx* c = L->ctx;
  if (!L->has_local) return SG_OK;
  std::vector<CopySeg> segs;
  std::vector<size_t> offs;
  size_t total = build_segs(f, true, segs, offs);
  SGCALL(ensure_dev_stage(c, total, segs.size(), true));
  CK(cudaStreamSynchronize(c->stream)); // staging buffers may still be in use by a previous transfer
  for (int b = 0; b < L->nbox; b++) {
    if (L->owner[b] != c->rank) continue;
    Box F = fab_rect(f, b, f->ng);
    REQUIRE(fabs[b], "sg_field_upload: null FAB pointer for owned box %d", b);
    memcpy(c->h_stage + offs[b], fabs[b], (size_t)F.nx() * F.ny() * f->ncomp * sizeof(double));
  }
  return scatter_from_stage(f, segs, c->h_stage, total);
}
// same, but the caller already holds all owned FABs consecutively (box order) in one (ideally pinned) buffer
extern "C" int sg_field_upload_packed(sg_field* f, const double* packed, size_t ndoubles) {
  REQUIRE(f && packed, "sg_field_upload_packed: bad arguments");
  sg_layout* L = f->lay;
  if (!L->has_local) return SG_OK;
  std::vector<CopySeg> segs;
  std::vector<size_t> offs;
  size_t total = build_segs(f, true, segs, offs);
  REQUIRE(total == ndoubles, "sg_field_upload_packed: buffer holds %zu doubles, the owned boxes need %zu", ndoubles, total);
  SGCALL(ensure_dev_stage(L->ctx, total, segs.size(), false));
  return scatter_from_stage(f, segs, packed, total);
}
static int gather_to_stage(const sg_field* f, std::vector<CopySeg>& segs, double* host_packed, size_t total) {
  sg_ctx* c = f->lay->ctx;
  CK(cudaMemcpyAsync(c->d_segs, segs.data(), segs.size() * sizeof(CopySeg), cudaMemcpyHostToDevice, c->stream));
  LAUNCH(c, k_copy_segs, dim3((unsigned)segs.size(), 4), 256, c->d_stage, f->base, c->d_segs, (int)segs.size());
  CK(cudaMemcpyAsync(host_packed, c->d_stage, total * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return SG_OK;
}
extern "C" int sg_field_download(const sg_field* f, double* const* fabs) {
  REQUIRE(f && fabs, "sg_field_download: bad arguments");
  sg_layout* L = f->lay;
  sg_ctx* c = L->ctx;
  if (!L->has_local) return SG_OK;
  std::vector<CopySeg> segs;
  std::vector<size_t> offs;
  size_t total = build_segs(f, false, segs, offs);
  SGCALL(ensure_dev_stage(c, total, segs.size(), true));
  SGCALL(gather_to_stage(f, segs, c->h_stage, total));
  for (int b = 0; b < L->nbox; b++) {
    if (L->owner[b] != c->rank) continue;
    Box F = fab_rect(f, b, f->ng);
    REQUIRE(fabs[b], "sg_field_download: null FAB pointer for owned box %d", b);
    memcpy(fabs[b], c->h_stage + offs[b], (size_t)F.nx() * F.ny() * f->ncomp * sizeof(double));
  }
  return SG_OK;
}
extern "C" int sg_field_download_packed(const sg_field* f, double* packed, size_t ndoubles) {
  REQUIRE(f && packed, "sg_field_download_packed: bad arguments");
  sg_layout* L = f->lay;
  if (!L->has_local) return SG_OK;
  std::vector<CopySeg> segs;
  std::vector<size_t> offs;
  size_t total = build_segs(f, false, segs, offs);
  REQUIRE(total == ndoubles, "sg_field_download_packed: buffer holds %zu doubles, the owned boxes need %zu", ndoubles, total);
  SGCALL(ensure_dev_stage(L->ctx, total, segs.size(), false));
  return gather_to_stage(f, segs, packed, total);
}
extern "C" int sg_field_device_view(sg_field* f, void** base, long long* pitch, long long* comp_stride, int patch_lo[2],
                                    int patch_hi[2], long long* offset_of_patch_lo) {
  REQUIRE(f && base, "sg_field_device_view");
  *base = f->base;
  if (pitch) *pitch = f->lay->pitch;
  if (comp_stride) *comp_stride = (long long)f->comp_stride;
  if (patch_lo) { patch_lo[0] = f->lay->patch.lo[0]; patch_lo[1] = f->lay->patch.lo[1]; }
  if (patch_hi) { patch_hi[0] = f->lay->patch.hi[0]; patch_hi[1] = f->lay->patch.hi[1]; }
  if (offset_of_patch_lo) *offset_of_patch_lo = (long long)SG_YOFF * f->lay->pitch + SG_XOFF;
  return SG_OK;
}

// ------------------------------------------------------------------------------------------------
// ghost filling
// ------------------------------------------------------------------------------------------------
// ghosts that come from other valid cells of the level: periodic images inside the patch, neighbouring ranks.
// depth <= 2.  Replaces LevelData::exchange (the in-patch box-to-box part needs no copy at all: boxes of one
// rank are merged into one array).
// Ghost cells on SK_GHOST sides of one-patch levels: periodic images inside the patch by a wrap kernel, rows owned by the
// neighbouring ranks by ncclSend/ncclRecv.  Several (field, depth) requests share ONE NCCL group (one fused transfer kernel per
// peer instead of one per field): the V-cycle is latency-bound on these exchanges at N > 1.
struct GhostReq { sg_field* f; int depth; };
// closes an NCCL group on every exit path (an error between group_start and group_end must not leave the group open)
struct NcclGroup {
  SgNccl& n; bool open = false;
  explicit NcclGroup(SgNccl& a) : n(a) {}
  int start() { int r = n.group_start(g_err); open = r == SG_OK; return r; }
  int end() { open = false; return n.group_end(g_err); }
  ~NcclGroup() { if (open) { std::string keep = g_err; n.group_end(g_err); g_err = keep; } }
};
static int fill_ghosts_multi(sg_ctx* c, const GhostReq* reqs, int n, cudaStream_t nccl_stream = nullptr) {
  if (!nccl_stream) nccl_stream = c->stream; // periodic wraps inside a patch always run on the main stream
  bool any_nccl = false;
  for (int q = 0; q < n; q++) {
    sg_field* f = reqs[q].f;
    const int depth = reqs[q].depth;
    sg_layout* L = f->lay;
    if (!L->has_local) continue;
    Geom g = make_geom(L, nullptr);
    int ex = f->cent == SG_XFACE, ey = f->cent == SG_YFACE;
    for (int comp = 0; comp < f->ncomp; comp++) {
      double* p = f->p(comp);
      if (L->wrap_local[0]) {
        int m = (g.ny + ey) * depth;
        LAUNCH(c, k_wrap_ghost, (m + 127) / 128, 128, p, g, 0, depth, ex, ey);
      }
      if (L->wrap_local[1]) {
        int m = (g.nx + 2 * depth + ex) * depth;
        LAUNCH(c, k_wrap_ghost, (m + 127) / 128, 128, p, g, 1, depth, ex, ey);
      }
    }
    any_nccl |= (L->nbr[2] >= 0 || L->nbr[3] >= 0);
  }
  if (!any_nccl) return SG_OK;
  NcclGroup grp(c->nccl);
  SGCALL(grp.start());
  for (int q = 0; q < n; q++) {
    sg_field* f = reqs[q].f;
    const int depth = reqs[q].depth;
    sg_layout* L = f->lay;
    if (!L->has_local || (L->nbr[2] < 0 && L->nbr[3] < 0)) continue;
    const int ny = L->ny, ey = f->cent == SG_YFACE;
    // rows are contiguous (x fastest, full pitch incl. ghost columns): send/recv straight from the field
    size_t cnt = (size_t)depth * L->pitch;
    for (int comp = 0; comp < f->ncomp; comp++) {
      double* p = f->p(comp) - SG_XOFF;
      // receives are posted high-side first so that, when both neighbours are the same rank (2 ranks, periodic),
      // the peer's [low send, high send] order pairs with [high recv, low recv] here
      if (L->nbr[3] >= 0) SGCALL(c->nccl.recv(p + (ptrdiff_t)(ny + ey) * L->pitch, cnt, L->nbr[3], nccl_stream, g_err));
      if (L->nbr[2] >= 0) SGCALL(c->nccl.recv(p + (ptrdiff_t)(-depth) * L->pitch, cnt, L->nbr[2], nccl_stream, g_err));
      if (L->nbr[2] >= 0) SGCALL(c->nccl.send(p + (ptrdiff_t)ey * L->pitch, cnt, L->nbr[2], nccl_stream, g_err));
      if (L->nbr[3] >= 0) SGCALL(c->nccl.send(p + (ptrdiff_t)(ny - depth) * L->pitch, cnt, L->nbr[3], nccl_stream, g_err));
    }
  }
  SGCALL(grp.end());
  return SG_OK;
}
static int fill_ghosts(sg_field* f, int depth) {
  GhostReq r{f, depth};
  return fill_ghosts_multi(f->lay->ctx, &r, 1);
}
static bool has_ghost_sides(const sg_layout* L) { return L->side_ghost[0] || L->side_ghost[1] || L->side_ghost[2] || L->side_ghost[3]; }

static int phys_bc(sg_field* f, const sg_bc* bc, const double dx[2], int homogeneous) {
  sg_layout* L = f->lay;
  if (!L->has_local) return SG_OK;
  Geom g = make_geom(L, bc);
  bool any = false;
  for (int s = 0; s < 4; s++) any |= (g.kind[s] == SK_PHYS_DIRI || g.kind[s] == SK_PHYS_NEUM);
  if (!any) return SG_OK;
  int n = std::max(g.nx, g.ny);
  LAUNCH(L->ctx, k_bc_ghost, dim3((n + 127) / 128, 4), 128, f->p(), g, dx[0], dx[1], homogeneous);
  return SG_OK;
}

static int exchange_g(sg_field* f, int depth, int corners);
static int phys_bc_any(sg_field* f, const sg_bc* bc, const double dx[2], int homogeneous);
static int extrap_any(sg_field* f, int copy_only);
extern "C" int sg_exchange(sg_field* f, int corners) {
  REQUIRE(f, "sg_exchange: null");
  // streaming-path layouts move rows/columns at full width, so corner ghosts are always consistent there
  if (!f->lay->fast) return f->lay->has_local ? exchange_g(f, std::max(1, std::min(f->ng, 2)), corners) : SG_OK;
  return fill_ghosts(f, std::max(1, std::min(f->ng, 2)));
}
extern "C" int sg_apply_bc(sg_field* f, const sg_bc* bc, const double dx[2], int homogeneous) {
  REQUIRE(f && bc && dx, "sg_apply_bc: null");
  return phys_bc_any(f, bc, dx, homogeneous);
}
static int extrap_ghost(sg_field* f, int copy_only) {
  sg_layout* L = f->lay;
  if (!L->has_local) return SG_OK;
  REQUIRE(f->cent == SG_CELL, "ExtrapGhostCells: cell data only");
  Geom g = make_geom(L, nullptr);
  for (int dir = 0; dir < 2; dir++) {
    int n = (dir == 0 ? g.ny : g.nx) + 2;
    LAUNCH(L->ctx, k_extrap_ghost, dim3((n + 127) / 128, f->ncomp), 128, f->p(), g, dir, copy_only, f->comp_stride, f->ncomp);
  }
  return SG_OK;
}
extern "C" int sg_extrap_ghost_cells(sg_field* f) { REQUIRE(f, "null"); return extrap_any(f, 0); }
extern "C" int sg_copy_ghost_cells(sg_field* f) { REQUIRE(f, "null"); return extrap_any(f, 1); }

#include "sg_general_host.inc"

// ------------------------------------------------------------------------------------------------
// stand-alone field kernels
// ------------------------------------------------------------------------------------------------
extern "C" int sg_nonlinear_level(const sg_params* p, sg_field* nl, sg_field* dnl, const sg_field* u, const sg_field* B,
                                  const sg_field* mask, const sg_field* Pi, const sg_field* zb) {
  REQUIRE(p && nl && dnl && u && B && mask && Pi && zb, "sg_nonlinear_level: null");
  sg_layout* L = u->lay;
  if (!L->has_local) return SG_OK;
  LAUNCH(L->ctx, k_nl_g, grid_g(L, 0, 0), B2D, nl->cb(), dnl->cb(), u->cb(), B->cb(), mask->cb(), Pi->cb(), zb->cb(), L->d_patches, phys(*p));
  return SG_OK;
}
static int gradient_cc_any(sg_field* grad2, sg_field* phi, const sg_field* mask, const double dx[2]) {
  sg_layout* L = phi->lay;
  if (!L->has_local) return SG_OK;
  LAUNCH(L->ctx, k_gradient_cc_g, grid_g(L, 0, 0), B2D, grad2->cb(0), grad2->cb(1), phi->cb(), mask ? mask->cb() : nullptr, L->d_patches, dx[0], dx[1]);
  return SG_OK;
}
extern "C" int sg_gradient_cc(sg_field* grad2, sg_field* phi, const sg_field* mask, const double dx[2]) {
  REQUIRE(grad2 && phi && dx && grad2->ncomp >= 2, "sg_gradient_cc: bad arguments");
  return gradient_cc_any(grad2, phi, mask, dx);
}
extern "C" int sg_compute_re(const sg_params* p, sg_field* Re, const sg_field* B, const sg_field* gradH) {
  REQUIRE(p && Re && B && gradH && gradH->ncomp >= 2, "sg_compute_re: bad arguments");
  sg_layout* L = Re->lay;
  if (!L->has_local) return SG_OK;
  LAUNCH(L->ctx, k_compute_re_g, grid_g(L, 2, 2), B2D, Re->cb(), B->cb(), gradH->cb(0), gradH->cb(1), L->d_patches, phys(*p));
  return SG_OK;
}
extern "C" int sg_divergence(sg_field* div, const sg_field* ux, const sg_field* uy, const double dx[2]) {
  REQUIRE(div && ux && uy && dx, "sg_divergence: null");
  sg_layout* L = div->lay;
  if (!L->has_local) return SG_OK;
  LAUNCH(L->ctx, k_divergence_g, grid_g(L, 0, 0), B2D, div->cb(), ux->cb(), uy->cb(), L->d_patches, dx[0], dx[1]);
  return SG_OK;
}

// AmrHydro::WFlx_level (src/AmrHydro.cpp:1415-1539).  With a coarser level (op->link): coarse gradient with dx*2 and the
// coarse level's ice mask, exchange + ExtrapGhostCells, QuadCFInterp of both components into the fine coarse-fine ghosts.
static int wflx_impl(sg_ctx* ctx, sg_op* op, const sg_params* p, sg_field* bX, sg_field* bY, sg_field* u, sg_field* u_coarse,
                     const sg_field* B, const sg_field* mask, const double dx[2]) {
  sg_layout* L = u->lay;
  sg_field *grad = nullptr, *Re = nullptr;
  int ngsave = 0;
  if (L->has_local) {
    SGCALL(ws_field(L, 1, 2, &grad));
    SGCALL(ws_field(L, 2, 1, &Re));
    SGCALL(gradient_cc_any(grad, u, p->use_mask_grad ? mask : nullptr, dx));
    ngsave = grad->ng;
    grad->ng = 1;
  }
  if (u_coarse) { // collective: every rank computes its part of the coarse gradient and takes part in the copy plan
    REQUIRE(op && op->link, "WFlx_level: a coarse head needs the operator's coarse-fine interpolator (use sg_op_UpdateOperator)");
    sg_layout* Lc = u_coarse->lay;
    sg_field* gradC = nullptr;
    if (Lc->has_local) {
      SGCALL(ws_field(Lc, 3, 2, &gradC));
      double dxc[2] = {dx[0] * 2, dx[1] * 2}; // "assumes refRatio = 2" (src/AmrHydro.cpp:1467)
      // (the coarse field's boundary ghost cells were filled by the caller / the coarse level's own UpdateOperator, as in the reference)
      if (Lc->fast && Lc->patches.size() == 1 && ctx->tune[9] != 1)
        SGCALL(coarse_gradient_near_fine(op, gradC, u_coarse, p->use_mask_grad ? op->link->crse_mask : nullptr, dxc));
      else SGCALL(gradient_cc_any(gradC, u_coarse, p->use_mask_grad ? op->link->crse_mask : nullptr, dxc));
      int ngc = gradC->ng;
      gradC->ng = 1;
      SGCALL(exchange_any(gradC, 1, 1));
      SGCALL(extrap_any(gradC, 0));
      gradC->ng = ngc;
    }
    sg_field dummyC, dummyF; // ranks without patches on one of the two levels still run the plan (with nothing to move there)
    dummyC.lay = Lc; dummyC.ncomp = 2; dummyC.ng = 1; dummyC.cent = SG_CELL; dummyC.base = nullptr; dummyC.comp_stride = 0;
    dummyF.lay = L; dummyF.ncomp = 2; dummyF.ng = 1; dummyF.cent = SG_CELL; dummyF.base = nullptr; dummyF.comp_stride = 0;
    SGCALL(cf_interp_impl(op, grad ? grad : &dummyF, gradC ? gradC : &dummyC));
  }
  if (!L->has_local) return SG_OK;
  SGCALL(exchange_any(grad, 1, 1)); // lvlgradH.exchange()
  SGCALL(extrap_any(grad, 0));      // ExtrapGhostCells(lvlgradH, levelDomain)
  grad->ng = ngsave;
  (void)Re; // Re over the ghosted box is evaluated inside the face kernel (never stored)
  LAUNCH(ctx, k_re_bcoef_g, dim3((L->max_nx + 1 + RB_TX - 1) / RB_TX, (L->max_ny + 1 + RB_TY - 1) / RB_TY, (unsigned)L->patches.size()), B2D, bX->cb(), bY->cb(), grad->cb(0), grad->cb(1), B->cb(), mask->cb(), L->d_patches, phys(*p),
         L->domain.lo[0], L->domain.lo[1], L->domain.hi[0], L->domain.hi[1]);
  return SG_OK;
}
extern "C" int sg_wflx_level(sg_ctx* ctx, const sg_params* p, sg_field* bX, sg_field* bY, sg_field* u, const sg_field* u_coarse,
                             const sg_field* B, const sg_field* mask, const double dx[2]) {
  REQUIRE(ctx && p && bX && bY && u && B && mask && dx, "sg_wflx_level: null");
  if (u_coarse) return fail(SG_ERR_UNSUPPORTED, "sg_wflx_level: with a coarse head the coarse-fine stencils of the level's operator "
                                                "are needed; call sg_op_UpdateOperator(op, phi, phi_coarse, ...)");
  return wflx_impl(ctx, nullptr, p, bX, bY, u, nullptr, B, mask, dx);
}

#include "sg_picard_host.inc"
#include "sg_regrid.inc"

// ------------------------------------------------------------------------------------------------
// factory
// ------------------------------------------------------------------------------------------------
extern "C" int sg_factory_define(sg_ctx* ctx, sg_factory** out, int nlevels, sg_layout* const* grids, const int* ref_ratios,
                                 const double coarse_dx[2], const sg_bc* bc, double alpha, sg_field* const* aCoef, double beta,
                                 sg_field* const* bCoefX, sg_field* const* bCoefY, const sg_params* params,
                                 sg_field* const* B, sg_field* const* Pi, sg_field* const* zb, sg_field* const* iceMask) {
  REQUIRE(ctx && out && nlevels >= 1 && grids && coarse_dx && bc && aCoef && bCoefX && bCoefY && params && B && Pi && zb && iceMask,
          "sg_factory_define: null argument");
  REQUIRE(nlevels == 1 || ref_ratios, "sg_factory_define: ref_ratios needed");
  sg_factory* f = new sg_factory();
  f->ctx = ctx; f->nlevels = nlevels;
  f->bc = *bc; f->alpha = alpha; f->beta = beta; f->prm = *params;
  f->dx.resize(nlevels);
  f->dx[0] = {coarse_dx[0], coarse_dx[1]};
  for (int l = 0; l < nlevels; l++) {
    f->grids.push_back(grids[l]);
    f->aCoef.push_back(aCoef[l]); f->bX.push_back(bCoefX[l]); f->bY.push_back(bCoefY[l]);
    f->B.push_back(B[l]); f->Pi.push_back(Pi[l]); f->zb.push_back(zb[l]); f->mask.push_back(iceMask[l]);
    f->ref_ratios.push_back(l < nlevels - 1 ? ref_ratios[l] : (ref_ratios ? ref_ratios[std::max(0, nlevels - 2)] : 2));
    if (l > 0) f->dx[l] = {f->dx[l - 1][0] / f->ref_ratios[l - 1], f->dx[l - 1][1] / f->ref_ratios[l - 1]};
  }
  *out = f;
  return SG_OK;
}
extern "C" int sg_factory_destroy(sg_factory* f) { delete f; return SG_OK; }
extern "C" int sg_factory_refToFiner(const sg_factory* f, int level, int* out) {
  REQUIRE(f && out, "null");
  if (level < 0 || level >= f->nlevels) return fail(SG_ERR_ABORT, "Domain not found in AMR hierarchy");
  *out = f->ref_ratios[level];
  return SG_OK;
}

// coefficient ghosts on SK_GHOST sides (periodic images / neighbouring ranks): depth 1, needed by the fused sweep
static int coef_ghost_depth(const sg_op* op) { return std::min(8, std::min(op->lay->nx, op->lay->ny)); }
static int coef_ghosts(sg_op* op, bool only_b) {
  if (!has_ghost_sides(op->lay)) return SG_OK;
  // depth 8: communication-avoiding relaxation updates up to 6 ghost rows (+ ring) on SK_GHOST sides; tiny levels get less
  const int d = coef_ghost_depth(op);
  GhostReq r[7] = {{op->bX, d}, {op->bY, d}, {op->B, d}, {op->Pi, d}, {op->zb, d}, {op->mask, d}, {op->aCoef, d}};
  return fill_ghosts_multi(op->ctx, r, only_b ? 2 : (op->alpha != 0.0 ? 7 : 6));
}

// The ice mask only enters the operator through COMPUTENONLINEARTERMS' test `mask < 0` (src/AmrHydroF.ChF:40).  Levels
// without a negative entry (every configuration but the valley geometry) need not stream the array at all; the scan runs
// when an operator is built or refreshed.  Multi-rank: decided per rank (a rank-local property of the rank-local cells,
// ghost cells included since the smoother recomputes the ring on them).
static int op_scan_mask(sg_op* op) {
  sg_layout* L = op->lay;
  op->mask_needed = true;
  if (!L->has_local) return SG_OK;
  sg_ctx* c = op->ctx;
  int* flag = reinterpret_cast<int*>(c->d_scalar + 120);
  CK(cudaMemsetAsync(flag, 0, sizeof(int), c->stream));
  if (L->fast) {
    // valid cells plus the ghost rows/columns the streaming sweeps may update (8 rows, 3 columns)
    LAUNCH(c, k_any_negative, grid2(L->nx + 6, L->ny + 16, B2D), B2D, op->mask->p() - 8 * (ptrdiff_t)L->pitch - 3, L->pitch, L->nx + 6, L->ny + 16, flag);
  } else {
    // patch table: the whole allocation (every patch with its ghost ring; padding is zero-filled at creation), as rows of 4096
    const long long n = (long long)op->mask->comp_stride;
    const int w = 4096, rows = (int)((n + w - 1) / w);
    if (rows > 1) LAUNCH(c, k_any_negative, grid2(w, rows - 1, B2D), B2D, op->mask->cb(), w, w, rows - 1, flag);
    const int tail = (int)(n - (long long)(rows - 1) * w);
    LAUNCH(c, k_any_negative, grid2(tail, 1, B2D), B2D, op->mask->cb() + (long long)(rows - 1) * w, w, tail, 1, flag);
  }
  int h = 1;
  CK(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  op->mask_needed = h != 0;
  return SG_OK;
}

static sg_op* new_op(sg_factory* f, int level) {
  sg_op* op = new sg_op();
  op->ctx = f->ctx;
  op->alpha = f->alpha; op->beta = f->beta; op->bc = f->bc; op->prm = f->prm;
  op->level = level;
  op->update_operator = f->prm.bcoeff_otf != 0;
  return op;
}

extern "C" int sg_op_destroy(sg_op* op) {
  if (!op) return SG_OK;
  if (op->owns_coefs) {
    sg_field_destroy(op->aCoef); sg_field_destroy(op->bX); sg_field_destroy(op->bY);
    sg_field_destroy(op->B); sg_field_destroy(op->Pi); sg_field_destroy(op->zb); sg_field_destroy(op->mask);
  }
  amr_link_free(op->link);
  fine_link_free(op->flink);
  if (op->owns_layout) sg_layout_destroy(op->lay);
  delete op;
  return SG_OK;
}

static int avg_cell(sg_field* c, const sg_field* f, int r) {
  sg_layout* Lc = c->lay;
  if (!Lc->has_local) return SG_OK;
  LAUNCH(Lc->ctx, k_avg_cell, grid2(Lc->nx, Lc->ny, B2D), B2D, c->p(), Lc->pitch, Lc->nx, Lc->ny, f->p(), f->lay->pitch, r);
  return SG_OK;
}
static int avg_face(sg_field* c, const sg_field* f, int r) {
  sg_layout* Lc = c->lay;
  if (!Lc->has_local) return SG_OK;
  int dir = c->cent == SG_XFACE ? 0 : 1;
  LAUNCH(Lc->ctx, k_avg_face, grid2(Lc->nx + 1, Lc->ny + 1, B2D), B2D, c->p(), Lc->pitch, Lc->nx, Lc->ny, f->p(), f->lay->pitch, r, dir);
  return SG_OK;
}

// arithmetic averages of the FINEST data by 2^depth (src/VCAMRNonLinearPoissonOp.cpp:1130-1139) + NeumBCForB (:1142-1148)
static int average_coefficients(sg_factory* f, sg_op* op, int level, int coarsening) {
  sg_layout* Lc = op->lay;
  if (f->alpha != 0.0) SGCALL(avg_cell(op->aCoef, f->aCoef[level], coarsening));
  SGCALL(avg_face(op->bX, f->bX[level], coarsening));
  SGCALL(avg_face(op->bY, f->bY[level], coarsening));
  SGCALL(avg_cell(op->B, f->B[level], coarsening));
  SGCALL(avg_cell(op->Pi, f->Pi[level], coarsening));
  SGCALL(avg_cell(op->zb, f->zb[level], coarsening));
  SGCALL(avg_cell(op->mask, f->mask[level], coarsening));
  if (Lc->has_local) {
    Geom g = make_geom(Lc, nullptr);
    int n = std::max(g.nx, g.ny);
    LAUNCH(f->ctx, k_neum_copy_ghost, dim3((n + 127) / 128, 4), 128, op->B->p(), g);
  }
  return SG_OK;
}

extern "C" int sg_factory_MGnewOp(sg_factory* f, int level, int depth, int homo_only, sg_op** out) {
  (void)homo_only;
  REQUIRE(f && out && level >= 0 && level < f->nlevels && depth >= 0, "sg_factory_MGnewOp: bad arguments");
  *out = nullptr;
  int coarsening = 1;
  for (int i = 0; i < depth; i++) coarsening *= 2;
  const int s_maxCoarse = 2; // src/AMRNonLinearPoissonOp.cpp:32
  if (coarsening > 1) {
    int ok = 0;
    SGCALL(sg_layout_coarsenable(f->grids[level], coarsening * s_maxCoarse, &ok));
    if (!ok) return SG_OK; // NULL: cannot coarsen further
  }
  sg_op* op = new_op(f, level);
  op->depth = depth;
  op->dx[0] = f->dx[level][0] * coarsening; op->dx[1] = f->dx[level][1] * coarsening;
  if (depth == 0) {
    op->lay = f->grids[level];
    op->aCoef = f->aCoef[level]; op->bX = f->bX[level]; op->bY = f->bY[level];
    op->B = f->B[level]; op->Pi = f->Pi[level]; op->zb = f->zb[level]; op->mask = f->mask[level];
  } else {
    sg_layout* Lc;
    int r = sg_layout_coarsen(f->grids[level], coarsening, &Lc);
    if (r != SG_OK) { delete op; return r; }
    op->lay = Lc; op->owns_layout = true; op->owns_coefs = true;
    SGCALL(sg_field_create(Lc, &op->aCoef, 1, 0, SG_CELL));
    SGCALL(sg_field_create(Lc, &op->bX, 1, 0, SG_XFACE));
    SGCALL(sg_field_create(Lc, &op->bY, 1, 0, SG_YFACE));
    SGCALL(sg_field_create(Lc, &op->B, 1, 1, SG_CELL));
    SGCALL(sg_field_create(Lc, &op->Pi, 1, 1, SG_CELL));
    SGCALL(sg_field_create(Lc, &op->zb, 1, 1, SG_CELL));
    SGCALL(sg_field_create(Lc, &op->mask, 1, 1, SG_CELL));
    SGCALL(average_coefficients(f, op, level, coarsening));
  }
  SGCALL(coef_ghosts(op, false));
  SGCALL(op_scan_mask(op));
  *out = op;
  return SG_OK;
}
extern "C" int sg_factory_AMRnewOp(sg_factory* f, int level, sg_op** out) {
  REQUIRE(f && out && level >= 0 && level < f->nlevels, "sg_factory_AMRnewOp: bad arguments");
  SGCALL(sg_factory_MGnewOp(f, level, 0, 0, out));
  // the four defines of src/VCAMRNonLinearPoissonOp.cpp:1215-1253 differ in which neighbours exist: a coarser level brings
  // the QuadCFInterp + coarsened-fine scratch; the flux register towards a finer level is built on first use (reflux)
  if (level > 0) {
    REQUIRE(f->ref_ratios[level - 1] == 2, "AMRnewOp: refinement ratio 2 only (WFlx_level assumes it, src/AmrHydro.cpp:1467)");
    int r = amr_link_build(*out, f->grids[level - 1], f->mask[level - 1]);
    if (r != SG_OK) { sg_op_destroy(*out); *out = nullptr; return r; }
  }
  return SG_OK;
}

// ------------------------------------------------------------------------------------------------
// operator
// ------------------------------------------------------------------------------------------------
static int check_same(const sg_op* op, const sg_field* f, const char* what) {
  if (!f) return fail(SG_ERR_INVALID, "%s: null field", what);
  const sg_layout *A = f->lay, *B = op->lay;
  if (A != B && (A->total != B->total || A->patches.size() != B->patches.size() || A->patch.lo[0] != B->patch.lo[0] ||
                 A->patch.lo[1] != B->patch.lo[1] || A->nx != B->nx || A->ny != B->ny))
    return fail(SG_ERR_INVALID, "%s: field layout differs from the operator's", what);
  return SG_OK;
}

// temporal blocking recomputes a 4-cell ring: periodic images inside the patch must then be at least 4 cells away
static bool two_iterations_fit(const sg_layout* L) {
  return (!L->wrap_local[0] || L->nx >= 4) && (!L->wrap_local[1] || L->ny >= 4) && L->ny >= 4;
}
// does relax_impl, in the default relax mode, run this level two iterations per launch (k_gsrb_twin)?  Levels the tile smoother
// does not take (more than 2 M cells: they stream from HBM) with enough rows to amortise a segment's nine load-only steps.
// tune key 18 = 1 turns it off (one iteration per launch, k_gsrb_stream).
static bool twin_smoother_applies(const sg_op* op) {
  const sg_layout* L = op->lay;
  const sg_ctx* c = op->ctx;
  if (!L->fast || c->relax_mode != 1 || c->tune[18] == 1 || !two_iterations_fit(L)) return false;
  if (c->tune[19] > 0) return (long long)L->nx * L->ny >= c->tune[19]; // tests: every level of at least that many cells
  // the variants that also stage the ice mask or aCoef hold 250 registers and measured slower than k_gsrb_stream (valley geometry,
  // 16384 x 4096: 0.91 against 0.88 ms per iteration): the default takes the plain variant only
  if (op->mask_needed || op->alpha != 0.0 || !op->prm.use_NL) return false;
  return (long long)L->nx * L->ny > (1LL << 21) && L->ny >= 64 && L->nx >= 64;
}
// does relax_impl run this level's sweeps four at a time in shared-memory tiles (k_gsrb_tile)?  One buffer swap per four iterations
// then, which the CUDA-graph eligibility test has to know (run_cycle)
static bool tile_smoother_applies(const sg_op* op) {
  const sg_layout* L = op->lay;
  const sg_ctx* c = op->ctx;
  if (!L->fast || c->relax_mode != 1 || c->tune[15] == 1 || (long long)L->nx * L->ny > (1LL << 21) || L->nx < 16 || L->ny < 16) return false;
  if (twin_smoother_applies(op)) return false;
  const Geom g = make_geom(L, &op->bc);
  const bool ygh = L->side_ghost[2] || L->side_ghost[3];
  if (ygh && (c->tune[3] != 0 || L->side_ghost[0] || L->side_ghost[1])) return false; // needs the 8-row exchange of the wide mode
  return g.kind[0] <= SK_PHYS_NEUM && g.kind[1] <= SK_PHYS_NEUM && (g.kind[2] <= SK_PHYS_NEUM || g.kind[2] == SK_GHOST) &&
         (g.kind[3] <= SK_PHYS_NEUM || g.kind[3] == SK_GHOST);
}
static int relax_swaps(const sg_op* op, int iterations) { // buffer swaps of one relax call
  if (tile_smoother_applies(op)) return iterations / 4 + iterations % 4;
  if (twin_smoother_applies(op) || ((op->ctx->relax_mode >= 3) && two_iterations_fit(op->lay))) return iterations / 2 + iterations % 2;
  return iterations;
}

// one levelGSRB iteration set
// trailing == false: the caller refills every ghost cell before its next read (the V-cycle driver does: restriction, residual
// and UpdateOperator all start with BC + exchange), so levelGSRB's closing exchange + homogeneous BC fill are dead stores
// phi_valid: depth to which the caller has just exchanged phi's ghost rows (0: unknown); the first sweep's exchange is skipped when
// that covers it
static int relax_impl(sg_op* op, sg_field* phi, const sg_field* rhs, int iterations, bool trailing = true, int phi_valid = 0) {
  NvtxRange nvtx_("levelGSRB");
  sg_layout* L = op->lay;
  sg_ctx* c = op->ctx;
  if (!L->has_local || iterations <= 0) return SG_OK;
  if (!L->fast) return relax_g(op, phi, rhs, iterations, trailing);
  OpArgs a = make_args(op);
  bool ghosts = has_ghost_sides(L);
  if (c->relax_mode >= 1) {
    // mode 1 (default): two iterations per sweep on HBM-sized levels (k_gsrb_twin), four per sweep in shared-memory tiles on
    // L2-resident ones (k_gsrb_tile), one iteration per sweep (k_gsrb_stream) for what remains; modes 3 / 4 / 5: two iterations
    // per sweep everywhere with k_gsrb_stream2 / k_gsrb_pair / k_gsrb_twin and a single sweep for an odd remainder; mode 2:
    // first-generation register-only sweep
    sg_field* scratch;
    SGCALL(ws_field(L, 0, 1, &scratch));
    const bool can2 = (c->relax_mode == 3 || c->relax_mode == 4 || c->relax_mode == 5) && two_iterations_fit(L);
    // default mode: levels too large for the tile smoother run two iterations per launch (k_gsrb_twin), one for an odd remainder
    const bool twin_default = twin_smoother_applies(op);
    // Communication-avoiding relaxation (mode 1): instead of exchanging two ghost rows before every sweep, exchange 2k rows
    // once per k <= 4 sweeps and let sweep s also update the ghost rows it still needs (2(k-1-s) per side) -- the same
    // arithmetic on the same values as their owner performs, so the result is unchanged while the NCCL (or wrap) calls drop
    // k-fold.  Needs ghost sides in y only and enough rows on both sides of the cut.
    const bool ygh_lo = L->side_ghost[2], ygh_hi = L->side_ghost[3];
    const bool wide = c->relax_mode == 1 && c->tune[3] == 0 && (ygh_lo || ygh_hi) && !L->side_ghost[0] && !L->side_ghost[1] && L->ny >= 16;
    bool rhs_pending = ghosts; // the right-hand side's ghost rows travel in the same NCCL group as the first exchange of phi
    const int rhs_depth = wide ? 7 : (can2 || twin_default) ? 3 : 1;
    FusedArgs f;
    f.a = a;
    f.rhs = rhs->p();
    f.sdx[0] = -op->dx[0]; f.sdx[1] = op->dx[0]; f.sdx[2] = -op->dx[1]; f.sdx[3] = op->dx[1];
    if (!c->smem_attr_set) {
      // warp-private staging rings: 4 warps x stages x (8|9) arrays x 32 lanes x 16 B
      CK(cudaFuncSetAttribute(k_gsrb_stream<0, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * GS_STAGES * 8 * 512));
      CK(cudaFuncSetAttribute(k_gsrb_stream<1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * GS_STAGES * 9 * 512));
      CK(cudaFuncSetAttribute(k_gsrb_stream2<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * GS2_STAGES * 8 * 512));
      CK(cudaFuncSetAttribute(k_gsrb_stream2<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * GS2_STAGES * 9 * 512));
      CK(cudaFuncSetAttribute(k_gsrb_pair<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (GP_STAGES * 8 * 512 + 1024)));
      CK(cudaFuncSetAttribute(k_gsrb_pair<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (GP_STAGES * 9 * 512 + 1024)));
      CK(cudaFuncSetAttribute((k_gsrb_twin<0, 0>), cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * TW_WARP_D2(5) * 16));
      CK(cudaFuncSetAttribute((k_gsrb_twin<0, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * TW_WARP_D2(6) * 16));
      CK(cudaFuncSetAttribute((k_gsrb_twin<1, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * TW_WARP_D2(7) * 16));
      c->smem_attr_set = true;
    }
    // Segments of rows per warp.  Measured on B200 (tools/relax_bench.py): many short segments beat one resident wave
    // (warps marching in lock-step), 32-64 rows per warp is the plateau on HBM-sized levels, and L2-resident levels want
    // the shortest segments that still amortise the load-only steps of a segment (4 for one iteration, 9 for two).
    auto plan = [&](int kind) { // 1 stream, 2 fused, 3 stream2, 4 producer/consumer pairs, 5 twin
      const int cols = kind == 3 ? GS2_COLS : kind == 4 ? GP_COLS : kind == 5 ? TW_COLS : kind == 1 ? GS_COLS : FUSED_COLS;
      f.nstrips = (L->nx + cols - 1) / cols;
      int minb = (kind == 5 || kind == 3) ? 2 : (kind == 1 || kind == 4) ? 3 : (c->tune[1] == 3 ? 3 : 4);
      const int wpc = 4;                                        // warps per CTA
      int capacity = c->num_sms * minb * (kind == 4 ? 2 : wpc); // resident warps (pairs for kind 4)
      int nsegs;
      const int nrows = (kind == 1 || kind == 5) ? f.yhi - f.ylo : L->ny;
      if (c->tune[0] > 0) nsegs = (nrows + c->tune[0] - 1) / c->tune[0];
      else if (kind == 1 || kind == 3 || kind == 4 || kind == 5) {
        long long rows = ((long long)f.nstrips * nrows) / (4LL * capacity);
        rows = kind != 1 ? std::max(16LL, std::min(96LL, rows)) : std::max((long long)(c->tune[15] > 0 ? c->tune[15] : 8), std::min(48LL, rows));
        nsegs = (int)((nrows + rows - 1) / rows);
      } else if (f.nstrips * ((L->ny + 63) / 64) <= capacity) nsegs = std::max(1, std::min((L->ny + 31) / 32, capacity / f.nstrips));
      else {
        nsegs = capacity / f.nstrips;               // one full wave ...
        if (nsegs < 1 || L->ny / nsegs > 1024) nsegs = (L->ny + 63) / 64; // ... unless segments get too long: many waves
      }
      f.rows_per_warp = (nrows + nsegs - 1) / nsegs;
      f.nsegs = (nrows + f.rows_per_warp - 1) / f.rows_per_warp;
      if (kind == 4) return (f.nstrips * f.nsegs + 1) / 2; // two pairs per CTA
      return (f.nstrips * f.nsegs + wpc - 1) / wpc;
    };
    f.ylo = 0; f.yhi = L->ny;
    int it = 0;
    // levels that sit in L2 (the coarser multigrid depths): four iterations per launch in shared-memory tiles (k_gsrb_tile) while at
    // least four remain.  x sides physical (Dirichlet / Neumann), y sides physical or ghost rows (then the eight rows of the
    // communication-avoiding exchange are exactly its halo).  tune key 15 = 1 turns it off.
    const bool tile_ok = tile_smoother_applies(op) && (!ghosts || wide);
    while (it < iterations) {
      if (tile_ok && iterations - it >= 4) {
        if (ghosts) {
          GhostReq r[2] = {{phi, 8}, {const_cast<sg_field*>(rhs), rhs_depth}};
          if (it == 0 && phi_valid >= 8) { if (rhs_pending) SGCALL(fill_ghosts_multi(c, r + 1, 1)); }
          else SGCALL(fill_ghosts_multi(c, r, rhs_pending ? 2 : 1));
          rhs_pending = false;
        }
        f.phi_in = phi->p();
        f.phi_out = scratch->p();
        dim3 tg((L->nx + GT_TX - 1) / GT_TX, (L->ny + GT_TY - 1) / GT_TY);
        const int nt = c->tune[14] >= 256 ? c->tune[14] : 1024; // threads per tile (tune key 14: 256 / 512 / 1024)
        if (a.has_a) { if (nt == 256) LAUNCH(c, (k_gsrb_tile<1, 4, 256>), tg, 256, f); else if (nt == 512) LAUNCH(c, (k_gsrb_tile<1, 4, 512>), tg, 512, f); else LAUNCH(c, (k_gsrb_tile<1, 4, 1024>), tg, 1024, f); }
        else { if (nt == 256) LAUNCH(c, (k_gsrb_tile<0, 4, 256>), tg, 256, f); else if (nt == 512) LAUNCH(c, (k_gsrb_tile<0, 4, 512>), tg, 512, f); else LAUNCH(c, (k_gsrb_tile<0, 4, 1024>), tg, 1024, f); }
        std::swap(phi->base, scratch->base);
        it += 4;
        continue;
      }
      const bool two = (can2 || twin_default) && it + 2 <= iterations;
      const int kind = two ? (c->relax_mode == 4 ? 4 : (c->relax_mode == 5 || twin_default) ? 5 : 3) : (c->relax_mode == 2 ? 2 : 1);
      const int ipl = two ? 2 : 1;                                                        // iterations per launch
      const int chunk = (wide && (kind == 1 || kind == 5)) ? std::min(4, iterations - it) / ipl : 1; // launches per ghost exchange
      // N > 1: the exchange of a chunk runs on the communication stream while the first sweep updates the rows that do not
      // depend on ghost rows ([m, ny-m) with m = 2 per iteration of a launch: a sweep over [a, b) reads rows [a-m, b+m)); the two
      // boundary strips follow once the ghost rows have landed.  Out of place, so the three launches write disjoint rows and read
      // the same input.
      auto launch_sweep = [&](int k) {
        const int blocks = plan(k);
        if (k == 5) {
          if (a.has_a) k_gsrb_twin<1, 1><<<blocks, 128, 4 * TW_WARP_D2(7) * 16, c->stream>>>(f);
          else if (a.use_mask || !a.prm.use_NL) k_gsrb_twin<0, 1><<<blocks, 128, 4 * TW_WARP_D2(6) * 16, c->stream>>>(f);
          else k_gsrb_twin<0, 0><<<blocks, 128, 4 * TW_WARP_D2(5) * 16, c->stream>>>(f);
        } else if (a.has_a) k_gsrb_stream<1, 3><<<blocks, 128, 4 * GS_STAGES * 9 * 512, c->stream>>>(f);
        else k_gsrb_stream<0, 3><<<blocks, 128, 4 * GS_STAGES * 8 * 512, c->stream>>>(f);
        c->launches++;
      };
      bool overlapped = false;
      if (ghosts) {
        const int need = 2 * ipl * chunk;
        GhostReq r[2] = {{phi, need}, {const_cast<sg_field*>(rhs), rhs_depth}};
        if (it == 0 && phi_valid >= need) { if (rhs_pending) SGCALL(fill_ghosts_multi(c, r + 1, 1)); }
        else if (wide && (kind == 1 || kind == 5) && c->comm_stream && c->tune[6] != 1 && L->ny >= 32 && (L->nbr[2] >= 0 || L->nbr[3] >= 0) &&
                 !L->wrap_local[0] && !L->wrap_local[1]) {
          CK(cudaEventRecord(c->ev_comm[0], c->stream));
          CK(cudaStreamWaitEvent(c->comm_stream, c->ev_comm[0], 0));
          SGCALL(fill_ghosts_multi(c, r, rhs_pending ? 2 : 1, c->comm_stream));
          CK(cudaEventRecord(c->ev_comm[1], c->comm_stream));
          overlapped = true;
        } else SGCALL(fill_ghosts_multi(c, r, rhs_pending ? 2 : 1));
        rhs_pending = false;
      }
      for (int sub = 0; sub < chunk; sub++) {
        const int ext = 2 * ipl * (chunk - 1 - sub); // ghost rows this launch still has to update for the launches after it
        f.phi_in = phi->p();
        f.phi_out = scratch->p();
        if (overlapped && sub == 0) {
          const int m = 2 * ipl;
          auto sweep_rows = [&](int ylo, int yhi) { f.ylo = ylo; f.yhi = yhi; launch_sweep(kind); };
          sweep_rows(m, L->ny - m);
          CK(cudaStreamWaitEvent(c->stream, c->ev_comm[1], 0));
          sweep_rows(ygh_lo ? -ext : 0, m);
          sweep_rows(L->ny - m, L->ny + (ygh_hi ? ext : 0));
          std::swap(phi->base, scratch->base);
          continue;
        }
        f.ylo = ygh_lo ? -ext : 0;
        f.yhi = L->ny + (ygh_hi ? ext : 0);
        if (kind == 4) {
          const int blocks = plan(kind);
          if (a.has_a) k_gsrb_pair<1><<<blocks, 128, 2 * (GP_STAGES * 9 * 512 + 1024), c->stream>>>(f);
          else k_gsrb_pair<0><<<blocks, 128, 2 * (GP_STAGES * 8 * 512 + 1024), c->stream>>>(f);
          c->launches++;
        } else if (kind == 5 || kind == 1) launch_sweep(kind);
        else if (kind == 3) {
          const int blocks = plan(kind);
          if (a.has_a) k_gsrb_stream2<1><<<blocks, 128, 4 * GS2_STAGES * 9 * 512, c->stream>>>(f);
          else k_gsrb_stream2<0><<<blocks, 128, 4 * GS2_STAGES * 8 * 512, c->stream>>>(f);
          c->launches++;
        } else {
          const int blocks = plan(kind);
          if (a.has_a) LAUNCH(c, (k_gsrb_fused<1, 4>), blocks, 128, f);
          else if (c->tune[1] == 3) LAUNCH(c, (k_gsrb_fused<0, 3>), blocks, 128, f);
          else LAUNCH(c, (k_gsrb_fused<0, 4>), blocks, 128, f);
        }
        std::swap(phi->base, scratch->base); // out-of-place sweep: the field now owns the new buffer
      }
      it += ipl * chunk;
    }
  } else {
    for (int it = 0; it < iterations; it++) {
      for (int pass = 0; pass < 2; pass++) {
        if (ghosts) SGCALL(fill_ghosts(phi, 1));
        SGCALL(phys_bc(phi, &op->bc, op->dx, 0));
        int half = (L->nx + 1) / 2;
        LAUNCH(c, k_gsrb_color, grid2(half, L->ny, B2D), B2D, phi->p(), rhs->p(), a, pass);
      }
    }
  }
  // trailing exchange + homogeneous BC fill (src/VCAMRNonLinearPoissonOp.cpp:751-759)
  if (!trailing) return SG_OK;
  if (ghosts) SGCALL(fill_ghosts(phi, 1));
  SGCALL(phys_bc(phi, &op->bc, op->dx, 1));
  return SG_OK;
}

extern "C" int sg_op_relax(sg_op* op, sg_field* phi, const sg_field* rhs, int iterations, int amr_fasmg_iter, int depth) {
  (void)amr_fasmg_iter; (void)depth;
  REQUIRE(op, "sg_op_relax: null op");
  SGCALL(check_same(op, phi, "relax(phi)"));
  SGCALL(check_same(op, rhs, "relax(rhs)"));
  REQUIRE(phi->ng >= 1, "relax: phi needs one ghost cell (CH_assert)");
  return relax_impl(op, phi, rhs, iterations);
}
extern "C" int sg_op_relaxNF(sg_op* op, sg_field* phi, const sg_field* phi_coarse, const sg_field* rhs, int iterations,
                             int amr_fasmg_iter, int depth, int print) {
  (void)print;
  REQUIRE(op && phi, "sg_op_relaxNF: null");
  if (phi_coarse) SGCALL(cf_interp_impl(op, phi, phi_coarse));
  return sg_op_relax(op, phi, rhs, iterations, amr_fasmg_iter, depth);
}

// BC -> exchange -> (NL fused) kernel; mode 0 apply, 1 residual, 2 residual + max-norm into d_scalar[slot]
static int apply_impl(sg_op* op, sg_field* out, sg_field* phi, const sg_field* rhs, int homogeneous, int mode, int slot, int ghost_depth = 1,
                      const unsigned char* special = nullptr) {
  sg_layout* L = op->lay;
  sg_ctx* c = op->ctx;
  if (!L->has_local) {
    // a rank without cells on this level still takes part in the allreduce of the norm: its contribution must be zero, not stale
    if (mode == 2 || mode == 3) CK(cudaMemsetAsync(reinterpret_cast<unsigned long long*>(c->d_scalar) + slot, 0, sizeof(double), c->stream));
    return SG_OK;
  }
  if (!L->fast) {
    if (mode >= 4) return fail(SG_ERR_UNSUPPORTED, "FAS coarse right-hand side accumulation and the fused composite sweeps exist on uniform (one-patch) levels only");
    return apply_g(op, out, phi, rhs, homogeneous, mode == 3 ? 2 : mode, slot, true);
  }
  OpArgs a = make_args(op);
  SGCALL(phys_bc(phi, &op->bc, op->dx, homogeneous));
  if (has_ghost_sides(L)) SGCALL(fill_ghosts(phi, ghost_depth));
  unsigned long long* nb = reinterpret_cast<unsigned long long*>(c->d_scalar) + slot;
  // 32 x 32 cells per block on large levels; tune key 4 = 1 forces the 32 x 8 shape, 100*bx + rows picks (bx, 256/bx) threads x rows
  // measured at 8192^2 (tools/apply_bench.py): 64 x 64 cells per block (16 rows per thread) 0.715 ms vs 0.80 ms for 32 x 8
  const bool tall = L->ny >= 256 && c->tune[4] != 1; // key 4 = 1: plain 32 x 8 blocks, one row per thread
  const bool big = tall && (long long)(L->nx / 64) * (L->ny / 64) >= 8LL * c->num_sms;
  int bx = big ? 64 : 32, rows = big ? 16 : tall ? 4 : 1;
  if (tall && c->tune[4] >= 100) { bx = c->tune[4] / 100; rows = c->tune[4] % 100; }
  dim3 blk(bx, 256 / bx);
  dim3 g((L->nx + bx - 1) / bx, (L->ny + blk.y * rows - 1) / (blk.y * rows));
  const size_t nblk = (size_t)g.x * g.y;
  unsigned long long* part = (mode == 2 || mode == 3 || mode == 6) && nblk <= SG_MAXPART && nblk >= 64 ? c->d_maxpart : nullptr;
#define APPLY_LAUNCH(M, ...)                                                     \
  do {                                                                           \
    if (rows == 1) LAUNCH(c, (k_apply<M, 1>), g, blk, __VA_ARGS__);              \
    else if (rows == 4) LAUNCH(c, (k_apply<M, 4>), g, blk, __VA_ARGS__);         \
    else if (rows == 8) LAUNCH(c, (k_apply<M, 8>), g, blk, __VA_ARGS__);         \
    else if (rows == 16) LAUNCH(c, (k_apply<M, 16>), g, blk, __VA_ARGS__);       \
    else if (rows == 32) LAUNCH(c, (k_apply<M, 32>), g, blk, __VA_ARGS__);       \
    else return fail(SG_ERR_INVALID, "tune key 4: rows per thread must be 1, 4, 8, 16 or 32"); \
  } while (0)
  if (mode == 0) APPLY_LAUNCH(0, out->p(), phi->p(), nullptr, a, nb);
  else if (mode == 1) APPLY_LAUNCH(1, out->p(), phi->p(), rhs->p(), a, nb);
  else if (mode == 4) APPLY_LAUNCH(4, out->p(), phi->p(), nullptr, a, nb);
  else if (mode == 5) APPLY_LAUNCH(5, out->p(), phi->p(), rhs->p(), a, nb);
  else if (mode == 6) APPLY_LAUNCH(6, nullptr, phi->p(), rhs->p(), a, nb, special + ((ptrdiff_t)SG_YOFF * L->pitch + SG_XOFF), part); // the map is addressed from the component base
  else {
    CK(cudaMemsetAsync(nb, 0, sizeof(double), c->stream));
    if (mode == 3) APPLY_LAUNCH(3, nullptr, phi->p(), rhs->p(), a, nb, nullptr, part);
    else APPLY_LAUNCH(2, out->p(), phi->p(), rhs->p(), a, nb, nullptr, part);
  }
  if (part) LAUNCH(c, k_max_partials, 1, 1024, part, nblk, nb);
#undef APPLY_LAUNCH
  return SG_OK;
}

extern "C" int sg_op_residual(sg_op* op, sg_field* lhs, sg_field* phi, const sg_field* rhs, int homogeneous) {
  REQUIRE(op, "sg_op_residual: null op");
  SGCALL(check_same(op, lhs, "residual(lhs)")); SGCALL(check_same(op, phi, "residual(phi)")); SGCALL(check_same(op, rhs, "residual(rhs)"));
  // AMRNonLinearPoissonOp::residual under FAS always calls residualI(..., false) (src/AMRNonLinearPoissonOp.cpp:247-253)
  (void)homogeneous;
  return apply_impl(op, lhs, phi, rhs, 0, 1, 0);
}
extern "C" int sg_op_residualNF(sg_op* op, sg_field* lhs, sg_field* phi, const sg_field* phi_coarse, const sg_field* rhs, int homogeneous) {
  REQUIRE(op, "sg_op_residualNF: null op");
  if (homogeneous) return fail(SG_ERR_ABORT, "VCAMRNonLinearPoissonOp::residualI homogeneous");
  SGCALL(check_same(op, lhs, "residualNF(lhs)")); SGCALL(check_same(op, phi, "residualNF(phi)")); SGCALL(check_same(op, rhs, "residualNF(rhs)"));
  if (phi_coarse) SGCALL(cf_interp_impl(op, phi, phi_coarse));
  return apply_impl(op, lhs, phi, rhs, 0, 1, 0);
}
extern "C" int sg_op_applyOp(sg_op* op, sg_field* lhs, sg_field* phi, int homogeneous) {
  REQUIRE(op, "sg_op_applyOp: null op");
  SGCALL(check_same(op, lhs, "applyOp(lhs)")); SGCALL(check_same(op, phi, "applyOp(phi)"));
  (void)homogeneous; // applyOp under FAS calls applyOpI(..., false) (src/AMRNonLinearPoissonOp.cpp:437-442)
  return apply_impl(op, lhs, phi, nullptr, 0, 0, 0);
}
extern "C" int sg_op_applyOpMg(sg_op* op, sg_field* lhs, sg_field* phi, sg_field* phi_coarse, int homogeneous) {
  REQUIRE(op, "sg_op_applyOpMg: null op");
  SGCALL(check_same(op, lhs, "applyOpMg(lhs)")); SGCALL(check_same(op, phi, "applyOpMg(phi)"));
  if (phi_coarse) SGCALL(cf_interp_impl(op, phi, phi_coarse)); // coarse domains of a single cell do not occur (block factor)
  return apply_impl(op, lhs, phi, nullptr, homogeneous, 0, 0); // applyOpI(lhs, phi, homogeneous)
}
extern "C" int sg_op_applyOpNoBoundary(sg_op* op, sg_field* lhs, sg_field* phi) {
  REQUIRE(op, "sg_op_applyOpNoBoundary: null op");
  SGCALL(check_same(op, lhs, "applyOpNoBoundary(lhs)")); SGCALL(check_same(op, phi, "applyOpNoBoundary(phi)"));
  sg_layout* L = op->lay;
  if (!L->has_local) return SG_OK;
  if (!L->fast) return apply_g(op, lhs, phi, nullptr, 0, 0, 0, false);
  OpArgs a = make_args(op);
  if (has_ghost_sides(L)) SGCALL(fill_ghosts(phi, 1));
  LAUNCH(op->ctx, (k_apply<0, 1>), grid2(L->nx, L->ny, B2D), B2D, lhs->p(), phi->p(), nullptr, a, nullptr);
  return SG_OK;
}

static int restrict_impl(sg_op* op, sg_field* resC, sg_field* phiC, sg_field* phiF, const sg_field* rhsF, sg_field* saveC = nullptr) {
  sg_layout* L = op->lay;
  sg_ctx* c = op->ctx;
  if (!L->has_local) return SG_OK;
  sg_layout* Lc = resC ? resC->lay : phiC->lay;
  if (!L->fast || !Lc->fast) return fail(SG_ERR_UNSUPPORTED, "restrictResidual/restrictR: multigrid descent exists below the base AMR level only "
                                                             "(levels above it use AMRRestrictS), and the base level is one patch per GPU");
  REQUIRE(Lc->nx * 2 == L->nx && Lc->ny * 2 == L->ny, "restrict: coarse layout is not the fine layout coarsened by 2");
  OpArgs a = make_args(op);
  if (resC) {
    SGCALL(phys_bc(phiF, &op->bc, op->dx, 0));
    if (has_ghost_sides(L)) SGCALL(fill_ghosts(phiF, 1));
    CK(cudaMemsetAsync(resC->base, 0, resC->comp_stride * sizeof(double), c->stream)); // res.setVal(0.0)
  }
  if (phiC) CK(cudaMemsetAsync(phiC->base, 0, phiC->comp_stride * sizeof(double), c->stream)); // phiCoarse.setVal(0.0)
  // large levels: 4 coarse rows per thread (tune key 5 = rows per thread: 1, 4, 8 or 16)
  const bool big = (long long)(Lc->nx / 32) * (Lc->ny / 64) >= 8LL * c->num_sms;
  const int rows = c->tune[5] > 0 ? c->tune[5] : big ? 4 : 1; // measured at 8192^2: 0.98 ms (1 row) -> 0.72 ms (4 rows), tools/apply_bench.py
  dim3 g((Lc->nx + 31) / 32, (Lc->ny + 8 * rows - 1) / (8 * rows));
#define RESTRICT_LAUNCH(R)                                                                                                         \
  do {                                                                                                                             \
    if (phiC) LAUNCH(c, (k_restrict<1, R>), g, B2D, resC ? resC->p() : nullptr, phiC->p(), saveC ? saveC->p() : nullptr, Lc->pitch, \
                     phiF->p(), rhsF ? rhsF->p() : nullptr, a);                                                                      \
    else LAUNCH(c, (k_restrict<0, R>), g, B2D, resC->p(), nullptr, nullptr, Lc->pitch, phiF->p(), rhsF->p(), a);                    \
  } while (0)
  if (rows == 1) RESTRICT_LAUNCH(1);
  else if (rows == 4) RESTRICT_LAUNCH(4);
  else if (rows == 8) RESTRICT_LAUNCH(8);
  else if (rows == 16) RESTRICT_LAUNCH(16);
  else return fail(SG_ERR_INVALID, "tune key 5: rows per thread must be 1, 4, 8 or 16");
#undef RESTRICT_LAUNCH
  return SG_OK;
}
extern "C" int sg_op_restrictResidual(sg_op* op, sg_field* res_coarse, sg_field* phi_fine, const sg_field* phi_coarse,
                                      const sg_field* rhs_fine, int homogeneous) {
  REQUIRE(op && res_coarse, "sg_op_restrictResidual: null");
  if (homogeneous) return fail(SG_ERR_ABORT, "VCAMRNonLinearPoissonOp::restrictResidual homogeneous");
  if (phi_coarse) SGCALL(cf_interp_impl(op, phi_fine, phi_coarse));
  SGCALL(check_same(op, phi_fine, "restrictResidual(phiFine)")); SGCALL(check_same(op, rhs_fine, "restrictResidual(rhsFine)"));
  return restrict_impl(op, res_coarse, nullptr, phi_fine, rhs_fine);
}
extern "C" int sg_op_restrictR(sg_op* op, sg_field* phi_coarse, const sg_field* phi_fine) {
  REQUIRE(op && phi_coarse, "sg_op_restrictR: null");
  SGCALL(check_same(op, phi_fine, "restrictR(phiFine)"));
  return restrict_impl(op, nullptr, phi_coarse, const_cast<sg_field*>(phi_fine), nullptr);
}
extern "C" int sg_op_prolongIncrement(sg_op* op, sg_field* phi, const sg_field* corr) {
  REQUIRE(op && corr, "sg_op_prolongIncrement: null");
  SGCALL(check_same(op, phi, "prolongIncrement(phi)"));
  sg_layout* L = op->lay;
  if (!L->has_local) return SG_OK;
  REQUIRE(L->fast && corr->lay->fast, "prolongIncrement: multigrid descent exists below the base AMR level only");
  REQUIRE(corr->lay->nx * 2 == L->nx && corr->lay->ny * 2 == L->ny, "prolongIncrement: coarse layout mismatch");
  LAUNCH(op->ctx, k_prolong, grid2(L->nx, L->ny, B2D), B2D, phi->p(), L->pitch, L->nx, L->ny, corr->p(), nullptr, corr->lay->pitch);
  return SG_OK;
}

extern "C" int sg_op_UpdateOperator(sg_op* op, sg_field* phi, const sg_field* phi_coarse, int depth, int amr_fasmg_iter, int homogeneous) {
  (void)depth; (void)amr_fasmg_iter;
  REQUIRE(op, "sg_op_UpdateOperator: null op");
  NvtxRange nvtx_("UpdateOperator");
  if (homogeneous) return fail(SG_ERR_ABORT, "VCAMRNonLinearPoissonOp::UpdateOperator homogeneous");
  SGCALL(check_same(op, phi, "UpdateOperator(phi)"));
  sg_layout* L = op->lay;
  if (!L->has_local && !phi_coarse) return SG_OK;
  if (L->has_local && L->fast && !phi_coarse && L->nx >= 2 && L->ny >= 2 && op->ctx->tune[12] == 1) {
    // tune key 12 = 1 -- one-patch level without a coarser one: gradient, its ghost cells, Re and both face coefficients in one pass
    // (k_update_op_fused).  The gradient's ghost rows towards a neighbouring GPU are computed from phi's second ghost row instead of
    // being exchanged: one halo exchange of phi (depth 2) replaces the two of phi and grad(phi).  Bit-exact, but NOT the default:
    // measured at 8192^2 on B200 (tools/updop_bench.py) 1.205 ms against 1.138 ms for the gradient + face kernels -- the pass is bound
    // by the FP64 sqrt/divide chain of COMPUTERE, which the fusion does not shorten, and the on-the-fly boundary logic adds to it.
    if (has_ghost_sides(L)) SGCALL(fill_ghosts(phi, 2));
    SGCALL(phys_bc(phi, &op->bc, op->dx, 0)); // the caller may read phi's boundary ghost cells afterwards, as after the reference's call
    OpArgs a = make_args(op);
    dim3 g((L->nx + 1 + RB_TX - 1) / RB_TX, (L->ny + 1 + RB_TY - 1) / RB_TY);
    if (op->prm.use_mask_grad) LAUNCH(op->ctx, k_update_op_fused<1>, g, B2D, op->bX->p(), op->bY->p(), phi->p(), a);
    else LAUNCH(op->ctx, k_update_op_fused<0>, g, B2D, op->bX->p(), op->bY->p(), phi->p(), a);
    return coef_ghosts(op, true);
  }
  if (L->has_local) {
    if (L->fast) { if (has_ghost_sides(L)) SGCALL(fill_ghosts(phi, 1)); }
    else SGCALL(exchange_g(phi, 1, 0));
    SGCALL(phys_bc_any(phi, &op->bc, op->dx, 0));
  }
  SGCALL(wflx_impl(op->ctx, op, &op->prm, op->bX, op->bY, phi, const_cast<sg_field*>(phi_coarse), op->B, op->mask, op->dx));
  if (L->has_local && L->fast) SGCALL(coef_ghosts(op, true));
  return SG_OK; // lambda is recomputed inside the kernels
}
extern "C" int sg_op_AverageOperator(sg_op* op, const sg_op* finest, int depth) {
  REQUIRE(op && finest && depth >= 0, "sg_op_AverageOperator: bad arguments");
  int coarsening = 1;
  for (int i = 0; i < depth; i++) coarsening *= 2;
  if (coarsening != 1) {
    SGCALL(avg_face(op->bX, finest->bX, coarsening));
    SGCALL(avg_face(op->bY, finest->bY, coarsening));
  }
  SGCALL(coef_ghosts(op, true));
  return SG_OK;
}
// does the smoother of this operator stream the ice mask?  (0: the level holds no negative mask entry, see op_scan_mask; the
// sweep then moves 64 instead of 72 bytes per cell-update -- what a roofline figure must be computed with)
extern "C" int sg_op_streams_mask(const sg_op* op, int* out) {
  REQUIRE(op && out, "sg_op_streams_mask: null");
  *out = op->mask_needed ? 1 : 0;
  return SG_OK;
}
// which kernel levelGSRB runs for this operator with the context's current relax mode and knobs: what a bench line names
extern "C" int sg_op_smoother_kind(const sg_op* op, int* out) {
  REQUIRE(op && out, "sg_op_smoother_kind: null");
  const sg_layout* L = op->lay;
  const sg_ctx* c = op->ctx;
  if (!L->fast) *out = (c->relax_mode != 0 && L->fused_state == 1 && c->tune[7] != 1) ? SG_SMOOTHER_PATCH : SG_SMOOTHER_COLOUR;
  else if (c->relax_mode == 0) *out = SG_SMOOTHER_COLOUR;
  else if (twin_smoother_applies(op) || (c->relax_mode == 5 && two_iterations_fit(L))) *out = SG_SMOOTHER_TWIN;
  else if (tile_smoother_applies(op)) *out = SG_SMOOTHER_TILE;
  else *out = SG_SMOOTHER_STREAM;
  return SG_OK;
}
extern "C" int sg_op_lambda(sg_op* op, sg_field* lam) {
  REQUIRE(op, "null");
  SGCALL(check_same(op, lam, "lambda"));
  sg_layout* L = op->lay;
  if (!L->has_local) return SG_OK;
  LAUNCH(op->ctx, k_lambda_g, grid_g(L, 0, 0), B2D, lam->cb(), L->d_patches, make_args_g(op));
  return SG_OK;
}
extern "C" int sg_op_createCoarser(sg_op* op, sg_field** coarse, const sg_field* fine, int ghosted) {
  (void)ghosted;
  REQUIRE(op && coarse && fine, "sg_op_createCoarser: null");
  int ok = 0;
  SGCALL(sg_layout_coarsenable(fine->lay, 2, &ok));
  REQUIRE(ok, "createCoarser: layout not coarsenable by 2 (CH_assert)");
  sg_layout* Lc;
  SGCALL(sg_layout_coarsen(fine->lay, 2, &Lc));
  int r = sg_field_create(Lc, coarse, fine->ncomp, fine->ng, fine->cent);
  if (r != SG_OK) { sg_layout_destroy(Lc); return r; }
  (*coarse)->owns_layout = true;
  return SG_OK;
}
extern "C" int sg_op_create(sg_op* op, sg_field** lhs, const sg_field* rhs) {
  REQUIRE(op && lhs && rhs, "sg_op_create: null");
  return sg_field_create(rhs->lay, lhs, rhs->ncomp, rhs->ng, rhs->cent);
}

// ---- vector surface -----------------------------------------------------------------------------
template <int OP>
static int vec_launch(sg_field* y, const sg_field* x, const sg_field* z, double a, double b, bool whole) {
  sg_layout* L = y->lay;
  if (!L->has_local) return SG_OK;
  if (!L->fast) return vec_g<OP>(y, x, z, a, b, whole);
  int ex = y->cent == SG_XFACE, ey = y->cent == SG_YFACE;
  int i0 = whole ? -y->ng : 0, i1 = L->nx + ex + (whole ? y->ng : 0), j0 = whole ? -y->ng : 0, j1 = L->ny + ey + (whole ? y->ng : 0);
  for (int c = 0; c < y->ncomp; c++)
    LAUNCH(L->ctx, k_vec<OP>, grid2(i1 - i0, j1 - j0, B2D), B2D, y->p(c), x ? x->p(c) : nullptr, z ? z->p(c) : nullptr, a, b, L->pitch, i0, i1, j0, j1);
  return SG_OK;
}
extern "C" int sg_op_assign(sg_op*, sg_field* lhs, const sg_field* rhs) { REQUIRE(lhs && rhs, "null"); return vec_launch<3>(lhs, rhs, nullptr, 0, 0, false); }
extern "C" int sg_op_assignLocal(sg_op*, sg_field* lhs, const sg_field* rhs) { REQUIRE(lhs && rhs, "null"); return vec_launch<3>(lhs, rhs, nullptr, 0, 0, true); }
extern "C" int sg_op_incr(sg_op*, sg_field* lhs, const sg_field* x, double scale) { REQUIRE(lhs && x, "null"); return vec_launch<1>(lhs, x, nullptr, scale, 0, false); }
extern "C" int sg_op_axby(sg_op*, sg_field* lhs, const sg_field* x, const sg_field* y, double a, double b) { REQUIRE(lhs && x && y, "null"); return vec_launch<0>(lhs, x, y, a, b, false); }
extern "C" int sg_op_scale(sg_op*, sg_field* lhs, double s) { REQUIRE(lhs, "null"); return vec_launch<2>(lhs, nullptr, nullptr, s, 0, false); }
extern "C" int sg_op_setToZero(sg_op*, sg_field* lhs) {
  REQUIRE(lhs, "null");
  if (!lhs->lay->has_local) return SG_OK;
  CK(cudaMemsetAsync(lhs->base, 0, lhs->comp_stride * lhs->ncomp * sizeof(double), lhs->lay->ctx->stream));
  return SG_OK;
}

// mode 0 max|x|, 1 sum|x|, 2 sum x^2, 3 sum x*y ; result (this rank) left in d_scalar[slot]
static int reduce_local(const sg_field* x, const sg_field* y, int mode, int slot) {
  sg_layout* L = x->lay;
  sg_ctx* c = L->ctx;
  CK(cudaMemsetAsync(c->d_scalar + slot, 0, sizeof(double), c->stream));
  if (!L->has_local) return SG_OK;
  if (!L->fast) return reduce_g(x, y, mode, slot);
  dim3 g = grid2(L->nx, L->ny, B2D);
  size_t nb = (size_t)g.x * g.y;
  if (mode != 0 && c->partial_cap < nb) {
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(c->d_partial);
    CK(cudaMalloc(&c->d_partial, nb * sizeof(double)));
    c->partial_cap = nb;
  }
  LAUNCH(c, k_reduce, g, B2D, x->p(), y ? y->p() : nullptr, L->pitch, L->nx, L->ny, mode, c->d_partial,
         reinterpret_cast<unsigned long long*>(c->d_scalar) + slot);
  if (mode != 0) LAUNCH(c, k_reduce_final, 1, 256, c->d_partial, (int)nb, c->d_scalar + slot);
  return SG_OK;
}
static int fetch_scalar(sg_ctx* c, int slot, double* out) {
  CK(cudaMemcpyAsync(c->h_scalar + slot, c->d_scalar + slot, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  *out = c->h_scalar[slot];
  return SG_OK;
}
static int global_reduce(sg_ctx* c, int slot, bool is_max) {
  if (c->nranks > 1) SGCALL(c->nccl.allreduce(c->d_scalar + slot, 1, is_max, c->stream, g_err));
  return SG_OK;
}
extern "C" int sg_op_norm(sg_op* op, const sg_field* x, int ord, double* out) {
  REQUIRE(op && x && out && ord >= 0 && ord <= 2, "sg_op_norm: bad arguments");
  REQUIRE(x->ncomp == 1, "norm: single-component fields only");
  SGCALL(reduce_local(x, nullptr, ord, 1));
  SGCALL(global_reduce(op->ctx, 1, ord == 0));
  SGCALL(fetch_scalar(op->ctx, 1, out));
  if (ord == 2) *out = std::sqrt(*out);
  return SG_OK;
}
extern "C" int sg_op_localMaxNorm(sg_op* op, const sg_field* x, double* out) {
  REQUIRE(op && x && out, "sg_op_localMaxNorm: null");
  SGCALL(reduce_local(x, nullptr, 0, 1));
  return fetch_scalar(op->ctx, 1, out);
}
extern "C" int sg_op_dotProduct(sg_op* op, const sg_field* a, const sg_field* b, double* out) {
  REQUIRE(op && a && b && out, "sg_op_dotProduct: null");
  SGCALL(reduce_local(a, b, 3, 1));
  SGCALL(global_reduce(op->ctx, 1, false));
  return fetch_scalar(op->ctx, 1, out);
}

// ---- AMR surface (src/AMRNonLinearPoissonOp.cpp:889-1264, VCAMRNonLinearPoissonOp.cpp:555-652) ------------------------
// AMROperator / NC / NF: coarse-fine interpolation (if a coarse phi is given), applyOpI, refluxing (if a fine phi is given)
static int amr_operator_impl(sg_op* op, sg_field* lof, sg_field* phiFine, sg_field* phi, const sg_field* phiCoarse, int hom, sg_op* fop) {
  if (phiCoarse) SGCALL(cf_interp_impl(op, phi, phiCoarse));
  SGCALL(apply_impl(op, lof, phi, nullptr, hom, 0, 0));
  if (phiFine) {
    REQUIRE(fop, "AMROperator: a fine phi needs the finer operator (CH_assert(a_finerOp != NULL))");
    SGCALL(reflux_impl(op, phiFine, phi, lof, fop));
  }
  return SG_OK;
}
static int amr_residual_impl(sg_op* op, sg_field* res, sg_field* phiFine, sg_field* phi, const sg_field* phiCoarse, const sg_field* rhs,
                             int hom, sg_op* fop) {
  if (!phiFine) { // AMRResidualNF: interpolate, then residualI (which aborts on homogeneous)
    if (hom) return fail(SG_ERR_ABORT, "VCAMRNonLinearPoissonOp::residualI homogeneous");
    if (phiCoarse) SGCALL(cf_interp_impl(op, phi, phiCoarse));
    return apply_impl(op, res, phi, rhs, 0, 1, 0);
  }
  SGCALL(amr_operator_impl(op, res, phiFine, phi, phiCoarse, hom, fop));
  return vec_launch<0>(res, res, rhs, -1.0, 1.0, false); // axby(residual, residual, rhs, -1, 1)
}
#define OPCHK(name) REQUIRE(op, name ": null op")
extern "C" int sg_op_AMRResidual(sg_op* op, sg_field* residual, const sg_field* phi_fine, sg_field* phi, const sg_field* phi_coarse,
                                 const sg_field* rhs, int hom, sg_op* finer_op) {
  OPCHK("AMRResidual");
  return amr_residual_impl(op, residual, const_cast<sg_field*>(phi_fine), phi, phi_coarse, rhs, hom, finer_op);
}
extern "C" int sg_op_AMRResidualNC(sg_op* op, sg_field* residual, const sg_field* phi_fine, sg_field* phi, const sg_field* rhs, int hom, sg_op* finer_op) {
  OPCHK("AMRResidualNC");
  SGCALL(amr_operator_impl(op, residual, const_cast<sg_field*>(phi_fine), phi, nullptr, hom, finer_op));
  return vec_launch<0>(residual, residual, rhs, -1.0, 1.0, false);
}
extern "C" int sg_op_AMRResidualNF(sg_op* op, sg_field* residual, sg_field* phi, const sg_field* phi_coarse, const sg_field* rhs, int hom) {
  OPCHK("AMRResidualNF");
  return amr_residual_impl(op, residual, nullptr, phi, phi_coarse, rhs, hom, nullptr);
}
extern "C" int sg_op_AMROperator(sg_op* op, sg_field* lofphi, const sg_field* phi_fine, sg_field* phi, const sg_field* phi_coarse, int hom, sg_op* finer_op) {
  OPCHK("AMROperator");
  return amr_operator_impl(op, lofphi, const_cast<sg_field*>(phi_fine), phi, phi_coarse, hom, finer_op);
}
extern "C" int sg_op_AMROperatorNC(sg_op* op, sg_field* lofphi, const sg_field* phi_fine, sg_field* phi, int hom, sg_op* finer_op) {
  OPCHK("AMROperatorNC");
  return amr_operator_impl(op, lofphi, const_cast<sg_field*>(phi_fine), phi, nullptr, hom, finer_op);
}
extern "C" int sg_op_AMROperatorNF(sg_op* op, sg_field* lofphi, sg_field* phi, const sg_field* phi_coarse, int hom) {
  OPCHK("AMROperatorNF");
  return amr_operator_impl(op, lofphi, nullptr, phi, phi_coarse, hom, nullptr);
}
// createCoarsened (src/AMRNonLinearPoissonOp.cpp:543-554): a field on this level's boxes coarsened by the ratio to the coarser level
extern "C" int sg_op_createCoarsened(sg_op* op, sg_field** out, const sg_field* fine, int ref_rat) {
  REQUIRE(op && out && fine && op->link, "createCoarsened: operator has no coarser level");
  REQUIRE(ref_rat == 2, "createCoarsened: refinement ratio 2 only");
  return sg_field_create(op->link->clay, out, fine->ncomp, fine->ng, SG_CELL);
}
// AMRRestrictS (:1027-1069): res_coarse lives on the coarsened-fine layout (createCoarsened)
static int amr_restrict_impl(sg_op* op, sg_field* resC, const sg_field* residual, sg_field* correction, const sg_field* coarseCorrection,
                             sg_field* scratch, int skip_res, bool keep_scratch = true) {
  REQUIRE(op->link, "AMRRestrictS: operator has no coarser level");
  sg_layout* L = op->lay;
  REQUIRE(resC->lay->patches.size() == L->patches.size(), "AMRRestrictS: res_coarse must live on the coarsened fine layout");
  const sg_field* from = scratch;
  if (!skip_res) SGCALL(amr_residual_impl(op, scratch, nullptr, correction, coarseCorrection, residual, 0, nullptr));
  else if (keep_scratch) SGCALL(vec_launch<3>(scratch, residual, nullptr, 0, 0, true)); // assignLocal
  else from = residual; // the V-cycle driver never reads the scratch copy: average straight from the field
  sg_layout* Lc = resC->lay;
  if (!L->has_local) return SG_OK;
  dim3 g((Lc->max_nx + B2D.x - 1) / B2D.x, (Lc->max_ny + B2D.y - 1) / B2D.y, (unsigned)Lc->patches.size());
  LAUNCH(op->ctx, k_amr_average, g, B2D, resC->cb(), Lc->d_patches, from->cb(), L->d_patches);
  return SG_OK;
}
extern "C" int sg_op_AMRRestrictS(sg_op* op, sg_field* res_coarse, const sg_field* residual, sg_field* correction,
                                  const sg_field* coarse_correction, sg_field* scratch, int skip_res) {
  REQUIRE(op && res_coarse && residual && scratch, "AMRRestrictS: null");
  return amr_restrict_impl(op, res_coarse, residual, correction, coarse_correction, scratch, skip_res);
}
// AMRProlongS / AMRProlongS_2 (:1105-1206): the coarsened-fine scratch and its copiers are the operator's own
// saved != nullptr (the V-cycle driver): coarseCorrection is the coarse level's NEW phi and `saved` its state before the coarse solve,
// copied onto the scratch's layout with the same plan; the correction phi_new - phi_saved (the driver's axby over the whole coarse
// level in the reference) is formed on the scratch, i.e. only where the prolongation reads it
static int amr_prolong_impl(sg_op* op, sg_field* correction, const sg_field* coarseCorrection, sg_op* crseOp, int second_order,
                            const sg_field* saved = nullptr) {
  AmrLink* K = op->link;
  REQUIRE(K, "AMRProlong: operator has no coarser level");
  sg_ctx* c = op->ctx;
  sg_layout* L = op->lay;
  if (second_order) {
    REQUIRE(crseOp, "AMRProlongS_2: needs the coarser operator");
    SGCALL(run_plan(c, K->temp->cb(), coarseCorrection->cb(), K->c2t[1]));
    if (saved && L->has_local) {
      sg_field one = *K->temp;
      one.ncomp = 1;
      const int ngs = one.ng;
      one.ng = 1;
      SGCALL(vec_launch<0>(&one, &one, saved, 1.0, -1.0, true)); // scratch = 1*phi_new + (-1)*phi_saved, ghost ring included
      one.ng = ngs;
    }
    if (L->has_local) {
      int ngs = K->temp->ng;
      K->temp->ng = 1;
      SGCALL(phys_bc_any(K->temp, &crseOp->bc, crseOp->dx, 0)); // m_use_FAS: inhomogeneous coarse BC on the scratch
      K->temp->ng = ngs;
      sg_field one = *K->temp;
      one.ncomp = 1;
      SGCALL(exchange_any(&one, 1, 1)); // CornerCopier exchange among the scratch's boxes
    }
  } else SGCALL(run_plan(c, K->temp->cb(), coarseCorrection->cb(), K->c2t[0]));
  if (L->has_local) LAUNCH(c, k_amr_prolong, grid_g(L, 0, 0), B2D, correction->cb(), L->d_patches, K->temp->cb(), K->clay->d_patches, second_order);
  return SG_OK;
}
extern "C" int sg_op_AMRProlongS(sg_op* op, sg_field* correction, const sg_field* coarse_correction) {
  REQUIRE(op && correction && coarse_correction, "AMRProlongS: null");
  return amr_prolong_impl(op, correction, coarse_correction, nullptr, 0);
}
extern "C" int sg_op_AMRProlongS_2(sg_op* op, sg_field* correction, const sg_field* coarse_correction, sg_op* coarse_op) {
  REQUIRE(op && correction && coarse_correction, "AMRProlongS_2: null");
  return amr_prolong_impl(op, correction, coarse_correction, coarse_op, 1);
}
extern "C" int sg_op_AMRUpdateResidual(sg_op* op, sg_field* residual, sg_field* correction, const sg_field* coarse_correction) {
  OPCHK("AMRUpdateResidual");
  return amr_residual_impl(op, residual, nullptr, correction, coarse_correction, residual, 0, nullptr);
}
extern "C" int sg_op_zeroCovered(sg_op* op, sg_field* coarse, const sg_field* fine_any) {
  REQUIRE(op && coarse && fine_any, "zeroCovered: null");
  return zero_covered_impl(op, coarse, fine_any->lay);
}
extern "C" int sg_op_AMRNorm(sg_op* op, const sg_field* coar, const sg_field* fine, int ref_rat, int ord, double* out) {
  OPCHK("AMRNorm");
  if (!fine) return sg_op_norm(op, coar, ord, out);
  REQUIRE(ref_rat == 2, "AMRNorm: refinement ratio 2 only");
  sg_field* tmp;
  SGCALL(ws_field(op->lay, 4, 1, &tmp));
  SGCALL(vec_launch<3>(tmp, coar, nullptr, 0, 0, false));
  SGCALL(zero_covered_impl(op, tmp, fine->lay));
  return sg_op_norm(op, tmp, ord, out);
}
extern "C" int sg_op_reflux(sg_op* op, const sg_field* phi_fine, const sg_field* phi, sg_field* residual, sg_op* finer_op) {
  REQUIRE(op && phi_fine && phi && residual && finer_op, "reflux: null");
  return reflux_impl(op, const_cast<sg_field*>(phi_fine), phi, residual, finer_op);
}
extern "C" int sg_op_cfInterp(sg_op* op, sg_field* phi, const sg_field* phi_coarse) {
  REQUIRE(op && phi && phi_coarse, "coarseFineInterp: null");
  return cf_interp_impl(op, phi, phi_coarse);
}
// LevelData::copyTo between two layouts of the same index space (valid cells of dst, grown by `ghosts` cells)
extern "C" int sg_field_copyTo(sg_field* dst, const sg_field* src, int ghosts) {
  REQUIRE(dst && src && ghosts >= 0 && ghosts <= 2, "sg_field_copyTo: bad arguments");
  return copy_to_impl(dst, src, ghosts);
}

// ---- virtuals the FAS path never calls but a Chombo-side subclass inherits ---------------------------------------
// AMRRestrict (src/AMRNonLinearPoissonOp.cpp:1011-1025): AMRRestrictS with a scratch created on the spot
extern "C" int sg_op_AMRRestrict(sg_op* op, sg_field* res_coarse, const sg_field* residual, sg_field* correction,
                                 const sg_field* coarse_correction, int skip_res) {
  REQUIRE(op && res_coarse && residual, "AMRRestrict: null");
  sg_field* scratch = nullptr;
  SGCALL(ws_field(op->lay, 5, 1, &scratch));
  return amr_restrict_impl(op, res_coarse, residual, correction, coarse_correction, scratch, skip_res);
}
// AMRProlong (:1073-1103): coarse correction copied onto the coarsened fine layout, PROLONGNL
extern "C" int sg_op_AMRProlong(sg_op* op, sg_field* correction, const sg_field* coarse_correction) {
  REQUIRE(op && correction && coarse_correction, "AMRProlong: null");
  return amr_prolong_impl(op, correction, coarse_correction, nullptr, 0);
}
// preCond, 2- and 3-argument forms (src/VCAMRNonLinearPoissonOp.cpp:174-208,233-271).  "Preconditioner is not used for FAS solve".
extern "C" int sg_op_preCond(sg_op* op, sg_field* phi, const sg_field* rhs) {
  REQUIRE(op, "preCond: null op");
  SGCALL(check_same(op, phi, "preCond(phi)")); SGCALL(check_same(op, rhs, "preCond(rhs)"));
  sg_layout* L = op->lay;
  if (L->has_local) LAUNCH(op->ctx, k_precond_init_g, grid_g(L, 0, 0), B2D, phi->cb(), rhs->cb(), L->d_patches, make_args_g(op));
  return relax_impl(op, phi, rhs, 2);
}
extern "C" int sg_op_preCond3(sg_op* op, sg_field* phi, const sg_field* res, const sg_field* rhs) {
  (void)res; // the initial guess from the residual is commented out in the reference (:254-265)
  REQUIRE(op, "preCond: null op");
  SGCALL(check_same(op, phi, "preCond(phi)")); SGCALL(check_same(op, rhs, "preCond(rhs)"));
  return relax_impl(op, phi, rhs, 2);
}
// getFlux, FluxBox form (src/VCAMRNonLinearPoissonOp.H:226-241 over VCAMRNonLinearPoissonOp.cpp:792-841): one direction per call
extern "C" int sg_op_getFlux(sg_op* op, sg_field* flux, const sg_field* phi, int dir, int ref, double scale) {
  REQUIRE(op && flux && phi && (dir == 0 || dir == 1), "getFlux: bad arguments (CH_assert(a_dir >= 0 && a_dir < SpaceDim))");
  REQUIRE(flux->cent == (dir == 0 ? SG_XFACE : SG_YFACE), "getFlux: flux must be face-centred in `dir` (CH_assert on the box type)");
  SGCALL(check_same(op, phi, "getFlux(phi)")); SGCALL(check_same(op, flux, "getFlux(flux)"));
  REQUIRE(phi->ng >= 1, "getFlux: phi needs a ghost cell (CH_assert(a_data.box().contains(ivlo)))");
  sg_layout* L = op->lay;
  if (!L->has_local) return SG_OK;
  const sg_field* bf = dir == 0 ? op->bX : op->bY;
  LAUNCH(op->ctx, k_get_flux_g, grid_g(L, dir == 0, dir == 1), B2D, flux->cb(), phi->cb(), bf->cb(), L->d_patches, dir, op->beta * ref / op->dx[dir], scale);
  return SG_OK;
}
// finerOperatorChanged (src/VCAMRNonLinearPoissonOp.cpp:1353-1438): multigrid coarsening of ALL operator data from `finer`
extern "C" int sg_op_finerOperatorChanged(sg_op* op, const sg_op* finer, int coarsening_factor) {
  REQUIRE(op && finer && coarsening_factor >= 1, "finerOperatorChanged: bad arguments");
  sg_layout *Lc = op->lay, *Lf = finer->lay;
  if (coarsening_factor != 1) {
    if (!Lc->fast || !Lf->fast) return fail(SG_ERR_UNSUPPORTED, "finerOperatorChanged: multigrid operators exist below the base AMR level only (one patch per GPU)");
    REQUIRE(op->owns_coefs, "finerOperatorChanged: this operator aliases the caller's coefficient fields (depth 0)");
    REQUIRE(Lc->nx * coarsening_factor == Lf->nx && Lc->ny * coarsening_factor == Lf->ny, "finerOperatorChanged: layouts do not differ by the coarsening factor");
    sg_field* cc[7] = {op->aCoef, op->B, op->Pi, op->zb, op->mask, op->bX, op->bY};
    const sg_field* cf[7] = {finer->aCoef, finer->B, finer->Pi, finer->zb, finer->mask, finer->bX, finer->bY};
    for (int k = 0; k < 7; k++) {
      if (Lc->has_local) CK(cudaMemsetAsync(cc[k]->base, 0, cc[k]->comp_stride * sizeof(double), op->ctx->stream)); // setVal(0.) on whole FABs
      if (k < 5) SGCALL(avg_cell(cc[k], cf[k], coarsening_factor));
      else SGCALL(avg_face(cc[k], cf[k], coarsening_factor));
    }
  }
  SGCALL(coef_ghosts(op, false)); // exchange(): box-to-box ghosts are the neighbours' cells in the merged storage
  return op_scan_mask(op);        // lambda is recomputed in the kernels; the mask-skip decision is what "needs resetting" here
}
// LevelDataOps::mDotProduct (src/AMRNonLinearPoissonOp.cpp:624-632)
extern "C" int sg_op_mDotProduct(sg_op* op, const sg_field* a, int n, const sg_field* const* b, double* out) {
  REQUIRE(op && a && n >= 0 && (n == 0 || (b && out)), "mDotProduct: bad arguments");
  for (int k = 0; k < n; k++) SGCALL(sg_op_dotProduct(op, a, b[k], out + k));
  return SG_OK;
}
// buildCopier / assignCopier (:577-597): Copier(rhs layout -> lhs layout, no ghost cells); the plan lives in the copy-plan cache
struct sg_copier { const sg_layout* src; const sg_layout* dst; };
extern "C" int sg_op_buildCopier(sg_op* op, sg_copier** out, const sg_field* lhs, const sg_field* rhs) {
  REQUIRE(op && out && lhs && rhs, "buildCopier: null");
  sg_copier* c = new sg_copier();
  c->src = rhs->lay; c->dst = lhs->lay;
  *out = c;
  return SG_OK;
}
extern "C" int sg_op_assignCopier(sg_op* op, sg_field* lhs, const sg_field* rhs, const sg_copier* copier) {
  REQUIRE(op && lhs && rhs && copier, "assignCopier: null");
  REQUIRE(copier->src == rhs->lay && copier->dst == lhs->lay, "assignCopier: copier was built for other layouts");
  return copy_to_impl(lhs, rhs, 0);
}
extern "C" int sg_copier_destroy(sg_copier* c) { delete c; return SG_OK; }
// setAlphaAndBeta / computeCoeffsOTF (src/VCAMRNonLinearPoissonOp.cpp:462-475)
extern "C" int sg_op_setAlphaAndBeta(sg_op* op, double alpha, double beta) {
  REQUIRE(op, "setAlphaAndBeta: null op");
  REQUIRE(alpha == 0.0 || op->aCoef, "setAlphaAndBeta: alpha != 0 needs an aCoef field");
  op->alpha = alpha; op->beta = beta; // lambda is never stored: nothing to reset ...
  return coef_ghosts(op, false);      // ... but with alpha != 0 the halo rows of aCoef on periodic / neighbour-GPU sides are now read
}
extern "C" int sg_op_computeCoeffsOTF(sg_op* op, int update_operator) {
  REQUIRE(op, "computeCoeffsOTF: null op");
  op->update_operator = update_operator != 0;
  return SG_OK;
}
// diagonalScale / divideByIdentityCoef (src/VCAMRNonLinearPoissonOp.H:157-175, "For TGA"): rhs *= aCoef, rhs /= aCoef
extern "C" int sg_op_diagonalScale(sg_op* op, sg_field* rhs, int kappa_weighted) {
  (void)kappa_weighted;
  REQUIRE(op && rhs && op->aCoef, "diagonalScale: null");
  SGCALL(check_same(op, rhs, "diagonalScale(rhs)"));
  return vec_launch<5>(rhs, op->aCoef, nullptr, 0, 0, false);
}
extern "C" int sg_op_divideByIdentityCoef(sg_op* op, sg_field* rhs) {
  REQUIRE(op && rhs && op->aCoef, "divideByIdentityCoef: null");
  SGCALL(check_same(op, rhs, "divideByIdentityCoef(rhs)"));
  return vec_launch<6>(rhs, op->aCoef, nullptr, 0, 0, false);
}
// homogeneousCFInterp (src/AMRNonLinearPoissonOp.cpp:1599-1795): dead under FAS (m_use_FAS hard-wired), kept for the surface
extern "C" int sg_op_homogeneousCFInterp(sg_op* op, sg_field* phi) {
  REQUIRE(op && phi, "homogeneousCFInterp: null");
  REQUIRE(phi->ng >= 1, "homogeneousCFInterp: phi needs one ghost cell (CH_assert)");
  AmrLink* K = op->link;
  if (!K) return SG_OK; // no coarser level: the CF region is empty
  SGCALL(check_same(op, phi, "homogeneousCFInterp(phi)"));
  CFHomo h;
  for (int d = 0; d < 2; d++) {
    const double dxf = op->dx[d], dxc = 2 * op->dx[d]; // m_dxCrse_vect = refRatio * dx
    h.c1[d] = 2 * (dxc - dxf) / (dxc + dxf); h.c2[d] = -(dxc - dxf) / (dxc + 3 * dxf); h.factor[d] = 1 - 2 * dxf / (dxf + dxc);
  }
  if (K->ncf)
    for (int comp = 0; comp < phi->ncomp; comp++)
      LAUNCH(op->ctx, k_cf_homogeneous, (K->ncf + 127) / 128, 128, phi->cb(comp), K->d_cf, K->ncf, h);
  return SG_OK;
}

// ------------------------------------------------------------------------------------------------
// FAS multigrid driver, device resident (absent fork's AMRFASMultiGrid / MultiGrid; see DESIGN.md for the
// inferred pieces, identical to oracle/suhmo_oracle.c)
// ------------------------------------------------------------------------------------------------
static void drop_solver_graphs(sg_ctx* c) {
  for (sg_solver* s : c->solvers)
    if (s->gexec) { cudaGraphExecDestroy(s->gexec); s->gexec = nullptr; s->gkey.clear(); }
}
extern "C" int sg_solver_define(sg_factory* f, sg_solver** out, int num_levels) {
  REQUIRE(f && out, "sg_solver_define: null");
  REQUIRE(num_levels >= 1 && num_levels <= f->nlevels, "sg_solver_define: %d levels requested, the factory holds %d", num_levels, f->nlevels);
  sg_solver* s = new sg_solver();
  s->ctx = f->ctx; s->fac = f; s->num_levels = num_levels;
  for (int depth = 0;; depth++) {
    sg_op* op = nullptr;
    int r = depth == 0 ? sg_factory_AMRnewOp(f, 0, &op) : sg_factory_MGnewOp(f, 0, depth, 1, &op);
    if (r != SG_OK) { return r; }
    if (!op) break;
    s->ops.push_back(op);
    sg_field *phi = nullptr, *rhs = nullptr, *save = nullptr, *tmp = nullptr;
    if (depth > 0) {
      SGCALL(sg_field_create(op->lay, &phi, 1, 1, SG_CELL));
      SGCALL(sg_field_create(op->lay, &rhs, 1, 0, SG_CELL));
      SGCALL(sg_field_create(op->lay, &save, 1, 1, SG_CELL));
      SGCALL(sg_field_create(op->lay, &tmp, 1, 0, SG_CELL));
    }
    s->phi.push_back(phi); s->rhs.push_back(rhs); s->save.push_back(save); s->tmp.push_back(tmp);
  }
  for (int l = 0; l < num_levels; l++) {
    sg_op* op = s->ops[0];
    if (l > 0) SGCALL(sg_factory_AMRnewOp(f, l, &op));
    s->aops.push_back(op);
    sg_field *res = nullptr, *corr = nullptr, *tmp = nullptr, *scr = nullptr, *resC = nullptr, *sav = nullptr;
    SGCALL(sg_field_create(op->lay, &res, 1, 0, SG_CELL));
    if (num_levels > 1) {
      SGCALL(sg_field_create(op->lay, &tmp, 1, 0, SG_CELL));
      SGCALL(sg_field_create(op->lay, &scr, 1, 1, SG_CELL));
      if (l > 0) { SGCALL(sg_field_create(op->link->clay, &resC, 1, 1, SG_CELL)); SGCALL(sg_field_create(op->link->clay, &sav, 1, 1, SG_CELL)); }
    }
    s->aresid.push_back(res); s->acorr.push_back(corr); s->atmp.push_back(tmp); s->ascratch.push_back(scr); s->aresC.push_back(resC);
    s->asave.push_back(sav);
  }
  s->resid = s->aresid[0];
  s->ctx->solvers.push_back(s);
  *out = s;
  return SG_OK;
}
// The reference builds a new factory + operator hierarchy for every Picard iteration (src/AmrHydro.cpp:704-735).
// Same effect without reallocating: re-average the (changed) finest coefficients into the existing MG operators.
extern "C" int sg_solver_refresh(sg_solver* s) {
  REQUIRE(s, "sg_solver_refresh: null");
  for (size_t d = 0; d < s->ops.size(); d++) {
    if (d > 0) SGCALL(average_coefficients(s->fac, s->ops[d], 0, 1 << d));
    SGCALL(coef_ghosts(s->ops[d], false));
    SGCALL(op_scan_mask(s->ops[d]));
  }
  return SG_OK;
}
// same, for the coefficients that change between the Picard iterations of one time step only (aCoef, bCoef)
extern "C" int sg_solver_refresh_bcoef(sg_solver* s) {
  REQUIRE(s, "sg_solver_refresh_bcoef: null");
  for (size_t d = 0; d < s->ops.size(); d++) {
    sg_op* op = s->ops[d];
    if (d > 0) {
      if (s->fac->alpha != 0.0) SGCALL(avg_cell(op->aCoef, s->fac->aCoef[0], 1 << d));
      SGCALL(avg_face(op->bX, s->fac->bX[0], 1 << d));
      SGCALL(avg_face(op->bY, s->fac->bY[0], 1 << d));
    }
    if (has_ghost_sides(op->lay)) {
      const int gd = coef_ghost_depth(op);
      GhostReq r[3] = {{op->bX, gd}, {op->bY, gd}, {op->aCoef, gd}};
      SGCALL(fill_ghosts_multi(op->ctx, r, op->alpha != 0.0 ? 3 : 2));
    }
  }
  return SG_OK;
}
extern "C" int sg_solver_destroy(sg_solver* s) {
  if (!s) return SG_OK;
  for (size_t d = 0; d < s->ops.size(); d++) {
    sg_field_destroy(s->phi[d]); sg_field_destroy(s->rhs[d]); sg_field_destroy(s->save[d]); sg_field_destroy(s->tmp[d]);
  }
  for (size_t l = 0; l < s->aops.size(); l++) {
    sg_field_destroy(s->aresid[l]); sg_field_destroy(s->acorr[l]); sg_field_destroy(s->atmp[l]); sg_field_destroy(s->ascratch[l]);
    sg_field_destroy(s->aresC[l]); sg_field_destroy(s->asave[l]);
    if (l > 0) sg_op_destroy(s->aops[l]);
  }
  for (sg_op* op : s->ops) sg_op_destroy(op);
  if (s->gexec) cudaGraphExecDestroy(s->gexec);
  auto& live = s->ctx->solvers;
  live.erase(std::remove(live.begin(), live.end(), s), live.end());
  delete s;
  return SG_OK;
}
extern "C" int sg_solver_depth(const sg_solver* s, int level, int* ndepth) {
  REQUIRE(s && ndepth && level >= 0 && level < s->num_levels, "sg_solver_depth");
  *ndepth = level == 0 ? (int)s->ops.size() : 1;
  return SG_OK;
}

static long long layout_cells_global(const sg_layout* L) {
  long long n = 0;
  for (const Box& b : L->boxes) n += b.npts();
  return n;
}
extern "C" int sg_solver_cell_updates_per_cycle(const sg_solver* s, const sg_solver_params* sp, double* out) {
  REQUIRE(s && sp && out, "null");
  double n = 0;
  int nd = (int)s->ops.size();
  for (int d = 0; d < nd; d++) n += (double)layout_cells_global(s->ops[d]->lay) * (d == nd - 1 ? sp->bottom : sp->pre + sp->post);
  for (int l = 1; l < s->num_levels; l++) n += (double)layout_cells_global(s->aops[l]->lay) * (sp->pre + sp->post);
  *out = n;
  return SG_OK;
}

// AverageOperator(ops[d], ops[0], d) for every depth d >= 1 at once: the finest face coefficients are read once per V-cycle.
// Same values as the per-depth calls (the depth-0 coefficients do not change inside a V-cycle), so mg_cycle only fills ghosts.
static int average_all_depths(sg_solver* s) {
  sg_op* op0 = s->ops[0];
  sg_layout* L = op0->lay;
  const int nd = (int)s->ops.size() - 1;
  s->coefs_averaged = false;
  if (nd < 1 || !L->fast) return SG_OK;
  for (int d = 1; d <= nd; d++) if (!s->ops[d]->lay->fast) return SG_OK;
  if (L->has_local) {
    for (int first = 1; first <= nd; first += 5) { // five depths per pass; deeper ones (ratio >= 64) take the per-depth kernel
      if (first > 1) {
        for (int d = first; d <= nd; d++) { SGCALL(avg_face(s->ops[d]->bX, op0->bX, 1 << d)); SGCALL(avg_face(s->ops[d]->bY, op0->bY, 1 << d)); }
        break;
      }
      AvgMulti ax, ay;
      ax.nd = ay.nd = std::min(5, nd);
      for (int k = 0; k < 5; k++) { ax.c[k] = ay.c[k] = nullptr; ax.pitch[k] = ay.pitch[k] = 0; }
      for (int k = 0; k < ax.nd; k++) {
        sg_op* o = s->ops[1 + k];
        ax.c[k] = o->bX->p(); ay.c[k] = o->bY->p();
        ax.pitch[k] = ay.pitch[k] = o->lay->pitch;
      }
      const int R = 1 << ax.nd;
      LAUNCH(s->ctx, k_avg_face_multi_x, dim3(((L->nx / 2 + 1) + 127) / 128, (L->ny + R - 1) / R), 128, op0->bX->p(), L->pitch, L->nx, L->ny, ax);
      LAUNCH(s->ctx, k_avg_face_multi_y, dim3((L->nx + 255) / 256, ((L->ny / 2 + 1) + AVGY_ROWS - 1) / AVGY_ROWS), 256, op0->bY->p(), L->pitch, L->nx, L->ny, ay);
    }
  }
  // coefficient ghost rows of every depth in one NCCL group (AverageOperator's coef_ghosts, batched)
  std::vector<GhostReq> reqs;
  for (int d = 1; d <= nd; d++)
    if (has_ghost_sides(s->ops[d]->lay)) {
      reqs.push_back({s->ops[d]->bX, coef_ghost_depth(s->ops[d])});
      reqs.push_back({s->ops[d]->bY, coef_ghost_depth(s->ops[d])});
    }
  if (!reqs.empty()) SGCALL(fill_ghosts_multi(s->ctx, reqs.data(), (int)reqs.size()));
  s->coefs_averaged = true;
  return SG_OK;
}
static int mg_cycle(sg_solver* s, int depth, sg_field* phi, sg_field* rhs, const sg_solver_params* sp, int phi_valid = 0) {
  NvtxRange nvtx_("MultiGrid::cycle");
  sg_op* op = s->ops[depth];
  int nd = (int)s->ops.size();
  if (depth == nd - 1) return relax_impl(op, phi, rhs, sp->bottom, false, phi_valid);
  SGCALL(relax_impl(op, phi, rhs, sp->pre, false, phi_valid));
  int dc = depth + 1;
  sg_op* opc = s->ops[dc];
  if (op->update_operator) {
    if (!s->coefs_averaged) SGCALL(sg_op_AverageOperator(opc, s->ops[0], dc)); // else: averaged and ghost-filled by average_all_depths
  }
  // restrictR + restrictResidual in one sweep over the fine level
  SGCALL(restrict_impl(op, s->rhs[dc], s->phi[dc], phi, rhs, s->save[dc]));              // ... and assignLocal(saved, phiC)
  // rhsC += applyOpMg(phiC, NULL, false); phiC's ghost rows are exchanged deep enough for the first sweeps of the coarse relax
  const int gd = std::min(8, std::min(opc->lay->nx, opc->lay->ny));
  SGCALL(apply_impl(opc, s->rhs[dc], s->phi[dc], nullptr, 0, 4, 0, gd));
  SGCALL(mg_cycle(s, dc, s->phi[dc], s->rhs[dc], sp, gd));
  if (op->lay->has_local) {                                                             // phi += I(phiC_new - phiC_saved)
    sg_layout* L = op->lay;
    LAUNCH(s->ctx, k_prolong, grid2(L->nx, L->ny, B2D), B2D, phi->p(), L->pitch, L->nx, L->ny, s->phi[dc]->p(), s->save[dc]->p(), opc->lay->pitch);
  }
  return relax_impl(op, phi, rhs, sp->post, false);
}
// AMRMultiGrid::computeAMRResidualLevel: aresid[l] = rhs[l] - L_composite(phi)
static int amr_residual_level(sg_solver* s, sg_field* const* phi, sg_field* const* rhs, int l, int l_max) {
  sg_field* fine = l < l_max ? phi[l + 1] : nullptr;
  const sg_field* crse = l > 0 ? phi[l - 1] : nullptr;
  return amr_residual_impl(s->aops[l], s->aresid[l], fine, phi[l], crse, rhs[l], 0, fine ? s->aops[l + 1] : nullptr);
}
// AMRFASMultiGrid::VCycle (absent fork; INFERRED, identical to oracle/suhmo_oracle.c:amr_vcycle -- see DESIGN.md)
static int amr_vcycle(sg_solver* s, sg_field* const* phi, sg_field* const* rhs, int ilev, int l_max, const sg_solver_params* sp, int iter) {
  NvtxRange nvtx_("AMRFASMultiGrid::VCycle level");
  sg_op* op = s->aops[ilev];
  if (op->update_operator) SGCALL(sg_op_UpdateOperator(op, phi[ilev], ilev > 0 ? phi[ilev - 1] : nullptr, ilev, iter, 0));
  if (ilev == 0) {
    if (op->update_operator) SGCALL(average_all_depths(s));
    int r = mg_cycle(s, 0, phi[0], s->aresid[0], sp);
    s->coefs_averaged = false;
    return r;
  }
  sg_op* opc = s->aops[ilev - 1];
  sg_ctx* c = s->ctx;
  // this level's right-hand side: the top level solves against the caller's rhs itself (the reference copies it into residual[lmax])
  sg_field* rl = ilev == l_max ? rhs[ilev] : s->aresid[ilev];
  // relaxNF = QuadCFInterp + levelGSRB; its closing exchange + homogeneous BC fill are dead stores here (see relax_impl)
  SGCALL(cf_interp_impl(op, phi[ilev], phi[ilev - 1]));
  SGCALL(relax_impl(op, phi[ilev], rl, sp->pre, false));
  // phi[ilev-1] <- average of phi[ilev] on the covered region (AMRRestrictS with skip_res), kept as the FAS reference state
  SGCALL(amr_restrict_impl(op, s->aresC[ilev], phi[ilev], phi[ilev], phi[ilev - 1], s->ascratch[ilev], 1, false));
  SGCALL(run_plan(c, phi[ilev - 1]->cb(), s->aresC[ilev]->cb(), op->link->t2c));
  // the reference state is only ever read back under and around level ilev (AMRProlongS_2's scratch): keep it there
  SGCALL(run_plan(c, s->asave[ilev]->cb(), phi[ilev - 1]->cb(), op->link->c2t[1]));
  // coarse right-hand side: composite residual, its covered part replaced by the averaged fine residual, plus AMROperatorNF(phi)
  const bool fused = opc->lay->fast && ilev - 1 == 0 && c->tune[9] != 1;
  if (fused) {
    // one sweep over the base level: (rhs - L phi) + L phi; the cells next to level ilev are refluxed and those under it redone below
    sg_field* res = s->aresid[ilev - 1];
    SGCALL(fine_link_build(opc, op->lay));
    SGCALL(apply_impl(opc, res, phi[ilev - 1], rhs[ilev - 1], 0, 5, 0));
    SGCALL(reflux_impl(opc, phi[ilev], phi[ilev - 1], res, op, 1, rhs[ilev - 1], 0));
    SGCALL(amr_restrict_impl(op, s->aresC[ilev], rl, phi[ilev], phi[ilev - 1], s->ascratch[ilev], 0));
    SGCALL(run_plan(c, res->cb(), s->aresC[ilev]->cb(), op->link->t2c));
    FineLink* F = opc->flink;
    if (F->nzero && opc->lay->has_local)
      LAUNCH(c, k_add_lof_segs, dim3((unsigned)F->nzero, 2), 128, res->cb(), phi[ilev - 1]->cb(), make_args(opc), opc->lay->patches[0].off, F->d_zero, F->nzero);
  } else {
    SGCALL(amr_residual_level(s, phi, rhs, ilev - 1, l_max));
    SGCALL(amr_restrict_impl(op, s->aresC[ilev], rl, phi[ilev], phi[ilev - 1], s->ascratch[ilev], 0));
    SGCALL(run_plan(c, s->aresid[ilev - 1]->cb(), s->aresC[ilev]->cb(), op->link->t2c));
    SGCALL(amr_operator_impl(opc, s->atmp[ilev - 1], nullptr, phi[ilev - 1], ilev - 1 > 0 ? phi[ilev - 2] : nullptr, 0, nullptr));
    SGCALL(vec_launch<1>(s->aresid[ilev - 1], s->atmp[ilev - 1], nullptr, 1.0, 0, false));
  }
  SGCALL(amr_vcycle(s, phi, rhs, ilev - 1, l_max, sp, iter));
  // AMRProlongS_2 of phi[ilev-1] - saved, with the operator's scratch standing in for m_resC
  SGCALL(amr_prolong_impl(op, phi[ilev], phi[ilev - 1], opc, 1, s->asave[ilev]));
  SGCALL(cf_interp_impl(op, phi[ilev], phi[ilev - 1]));
  return relax_impl(op, phi[ilev], rl, sp->post, false);
}
static int vcycle(sg_solver* s, sg_field* const* phi, sg_field* const* rhs, int l_max, const sg_solver_params* sp, int iter) {
  if (l_max == 0) {
    sg_op* op0 = s->ops[0];
    if (op0->update_operator) { SGCALL(sg_op_UpdateOperator(op0, phi[0], nullptr, 0, iter, 0)); SGCALL(average_all_depths(s)); }
    int r = mg_cycle(s, 0, phi[0], rhs[0], sp);
    s->coefs_averaged = false;
    return r;
  }
  return amr_vcycle(s, phi, rhs, l_max, l_max, sp, iter); // residual[lmax] = rhs[lmax]: aliased, not copied
}
// computeAMRResidual: max over levels of the max-norm of the composite residual (covered cells zeroed), left in
// d_scalar[slot] on all ranks
static int residual_norm(sg_solver* s, sg_field* const* phi, sg_field* const* rhs, int l_max, int slot) {
  NvtxRange nvtx_("computeAMRResidual + norm");
  if (l_max == 0) {
    SGCALL(apply_impl(s->ops[0], s->resid, phi[0], rhs[0], 0, 3, slot)); // max-norm only, the residual itself is not needed
    return global_reduce(s->ctx, slot, true);
  }
  sg_ctx* c = s->ctx;
  unsigned long long* nb = reinterpret_cast<unsigned long long*>(c->d_scalar) + slot;
  CK(cudaMemsetAsync(nb, 0, sizeof(double), c->stream));
  for (int l = l_max; l >= 0; l--) {
    sg_op* opl = s->aops[l];
    if (l == l_max && c->tune[9] != 1) {
      // finest level: nothing to reflux or zero, the max-norm rides on the residual sweep
      if (l > 0) SGCALL(cf_interp_impl(opl, phi[l], phi[l - 1]));
      SGCALL(apply_impl(opl, s->aresid[l], phi[l], rhs[l], 0, 2, slot));
      continue;
    }
    if (l == 0 && opl->lay->fast && c->tune[9] != 1) {
      // base level under a finer one: norm-only sweep over the cells away from the finer level, register cells by the sparse kernel,
      // covered cells count as zero
      SGCALL(fine_link_build(opl, s->aops[1]->lay));
      if (opl->lay->has_local) SGCALL(apply_impl(opl, nullptr, phi[0], rhs[0], 0, 6, slot, 1, opl->flink->d_special));
      SGCALL(reflux_impl(opl, phi[1], phi[0], nullptr, s->aops[1], 2, rhs[0], slot));
      continue;
    }
    SGCALL(amr_residual_level(s, phi, rhs, l, l_max));
    if (l < l_max) SGCALL(zero_covered_impl(s->aops[l], s->aresid[l], s->aops[l + 1]->lay));
    // max|.| accumulates into the same slot across levels (atomicMax on the bit pattern)
    sg_layout* L = s->aops[l]->lay;
    if (!L->has_local) continue;
    if (L->fast) LAUNCH(c, k_reduce, grid2(L->nx, L->ny, B2D), B2D, s->aresid[l]->p(), nullptr, L->pitch, L->nx, L->ny, 0, c->d_partial, nb);
    else LAUNCH(c, k_reduce_g, grid_g(L, 0, 0), B2D, s->aresid[l]->cb(), nullptr, L->d_patches, 0, c->d_partial, nb);
  }
  return global_reduce(s->ctx, slot, true);
}

// One V-cycle followed by the residual norm (left in d_scalar[normslot]).  After one ordinary (warm-up) cycle with a given
// set of fields and parameters, the cycle is captured into a CUDA graph and replayed: valid because every array a kernel
// names is unchanged from cycle to cycle -- the out-of-place smoother swaps a field with its scratch buffer, and an even
// number of sweeps per relax call (pre, post and bottom all even, as in every SUHMO configuration) swaps it back.
static int run_cycle(sg_solver* s, sg_field* const* phi, sg_field* const* rhs, int l_max, const sg_solver_params* sp, int iter, int normslot) {
  sg_ctx* c = s->ctx;
  const int FIX = 100; // scratch slot the captured norm lands in
  // every field must own the same buffer after a cycle as before it: an even number of out-of-place launches per field
  bool even = c->relax_mode == 0;
  if (!even) {
    even = true;
    const int nd = (int)s->ops.size();
    for (int d = 0; d < nd; d++) {
      const int sw = d == nd - 1 ? relax_swaps(s->ops[d], sp->bottom) : relax_swaps(s->ops[d], sp->pre) + relax_swaps(s->ops[d], sp->post);
      if (sw % 2) even = false;
    }
    if (l_max > 0 && (sp->pre + sp->post) % 2) even = false; // refined levels: one swap per iteration
  }
  // N > 1: the cycle holds NCCL send/recv groups and the cross-stream events of the overlapped halo exchange; both are captured (every
  // rank issues the same sequence; a rank whose capture fails falls back to eager launches, which issue the same NCCL calls in the same
  // order).  Measured at N = 2: 23.97 -> 23.64 ms per 3-level V-cycle, same residual history.  tune key 13 = 2 turns it off.
  const bool eligible = (c->nranks == 1 || c->tune[13] != 2) && even && c->tune[2] == 0;
  if (!eligible) {
    SGCALL(vcycle(s, phi, rhs, l_max, sp, iter));
    return residual_norm(s, phi, rhs, l_max, normslot);
  }
  // Everything a captured kernel names or a launch shape depends on: parameters, every tuning knob, the fields, and -- because the
  // out-of-place smoother swaps a field's buffer with the layout's scratch buffer (ws slot 0) -- the scratch buffers of every
  // level and depth: a public relax with an odd sweep count on ANOTHER field of the same layout hands that field the scratch the
  // graph would still write to.
  std::vector<long long> key = {l_max, sp->pre, sp->post, sp->bottom, c->relax_mode, (long long)(size_t)c->stream};
  for (int k = 0; k < 32; k++) key.push_back(c->tune[k]);
  auto scratch_of = [](const sg_layout* L) { return (long long)(size_t)((!L->ws.empty() && L->ws[0]) ? L->ws[0]->base : nullptr); };
  for (int l = 0; l <= l_max; l++) { key.push_back((long long)(size_t)phi[l]->base); key.push_back((long long)(size_t)rhs[l]->base); }
  for (sg_op* op : s->aops) {
    key.push_back((long long)(size_t)op->bX->base); key.push_back((long long)(size_t)op->bY->base); key.push_back((long long)(size_t)op->B->base);
    key.push_back((long long)(size_t)op->Pi->base); key.push_back((long long)(size_t)op->zb->base); key.push_back((long long)(size_t)op->mask->base);
    key.push_back(op->mask_needed); key.push_back(scratch_of(op->lay));
  }
  for (size_t d = 0; d < s->ops.size(); d++) {
    key.push_back(s->ops[d]->mask_needed); key.push_back(scratch_of(s->ops[d]->lay));
    if (s->phi[d]) key.push_back((long long)(size_t)s->phi[d]->base);
  }
  if (s->gexec && key == s->gkey) {
    CK(cudaGraphLaunch(s->gexec, c->stream));
    c->launches += s->glaunches;
  } else if (key == s->warm_key) {
    if (s->gexec) { cudaGraphExecDestroy(s->gexec); s->gexec = nullptr; }
    long long l0 = c->launches;
    cudaGraph_t g = nullptr;
    // buffer ownership before the capture: an aborted capture may stop between two swaps of the out-of-place smoother
    std::vector<std::pair<sg_field*, double*>> saved;
    auto remember = [&](sg_field* f) { if (f) saved.push_back({f, f->base}); };
    for (int l = 0; l <= l_max; l++) { remember(phi[l]); if (!s->aops[l]->lay->ws.empty()) remember(s->aops[l]->lay->ws[0]); }
    for (size_t d = 0; d < s->ops.size(); d++) { remember(s->phi[d]); if (!s->ops[d]->lay->ws.empty()) remember(s->ops[d]->lay->ws[0]); }
    CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    int r = vcycle(s, phi, rhs, l_max, sp, iter);
    if (r == SG_OK) r = residual_norm(s, phi, rhs, l_max, FIX);
    cudaError_t e = cudaStreamEndCapture(c->stream, &g);
    if (r != SG_OK || e != cudaSuccess || !g) { // something in the cycle cannot be captured: run it the ordinary way from now on
      if (g) cudaGraphDestroy(g);
      cudaGetLastError();
      for (auto& kv : saved) kv.first->base = kv.second; // nothing captured has run: the buffers are as they were
      c->tune[2] = 1;
      c->launches = l0;
      SGCALL(vcycle(s, phi, rhs, l_max, sp, iter));
      return residual_norm(s, phi, rhs, l_max, normslot);
    }
    s->glaunches = c->launches - l0;
    CK(cudaGraphInstantiate(&s->gexec, g, 0));
    cudaGraphDestroy(g);
    s->gkey = key;
    CK(cudaGraphLaunch(s->gexec, c->stream));
  } else {
    s->warm_key = key;
    SGCALL(vcycle(s, phi, rhs, l_max, sp, iter));
    return residual_norm(s, phi, rhs, l_max, normslot);
  }
  if (normslot != FIX) CK(cudaMemcpyAsync(c->d_scalar + normslot, c->d_scalar + FIX, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  return SG_OK;
}

extern "C" int sg_solver_solve(sg_solver* s, sg_field* const* phi, sg_field* const* rhs, int l_max, int l_base,
                               const sg_solver_params* sp, double* hist, sg_solve_stats* stats) {
  REQUIRE(s && phi && rhs && sp && phi[0] && rhs[0], "sg_solver_solve: null");
  NvtxRange nvtx_("SolveForHead_nl");
  REQUIRE(l_base == 0 && l_max >= 0 && l_max < s->num_levels, "sg_solver_solve: l_base must be 0 and l_max below the number of levels defined");
  sg_ctx* c = s->ctx;
  for (int l = 0; l <= l_max; l++) {
    REQUIRE(phi[l] && rhs[l], "sg_solver_solve: null field on level %d", l);
    SGCALL(check_same(s->aops[l], phi[l], "solve(phi)")); SGCALL(check_same(s->aops[l], rhs[l], "solve(rhs)"));
  }
  long long l0 = c->launches;
  struct Events { // destroyed on every exit path
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~Events() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
  } evs;
  CK(cudaEventCreate(&evs.e0)); CK(cudaEventCreate(&evs.e1));
  cudaEvent_t e0 = evs.e0, e1 = evs.e1;
  double initial = 0, rnorm = 0;
  SGCALL(residual_norm(s, phi, rhs, l_max, 2));
  SGCALL(fetch_scalar(c, 2, &initial));
  rnorm = initial;
  if (hist) hist[0] = initial;
  int iter = 0, exit_status = 0;
  CK(cudaEventRecord(e0, c->stream));
  if (sp->fixed_cycles > 0) {
    REQUIRE(sp->fixed_cycles <= 60, "fixed_cycles <= 60");
    for (iter = 0; iter < sp->fixed_cycles; iter++) {
      SGCALL(run_cycle(s, phi, rhs, l_max, sp, iter, 3 + iter)); // norms stay on the device until the end
    }
    CK(cudaEventRecord(e1, c->stream));
    CK(cudaMemcpyAsync(c->h_scalar + 3, c->d_scalar + 3, sizeof(double) * sp->fixed_cycles, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < sp->fixed_cycles; k++) {
      if (hist) hist[k + 1] = c->h_scalar[3 + k];
      rnorm = c->h_scalar[3 + k];
    }
  } else {
    double norm_last = 2 * initial;
    bool goNorm = rnorm > sp->norm_thresh, goRedu = rnorm > sp->eps * initial, goIter = iter < sp->max_iter;
    bool goHang = iter < sp->imin || rnorm < (1 - sp->hang) * norm_last, goMin = iter < sp->iter_min;
    while (goMin || (goIter && goRedu && goHang && goNorm)) {
      norm_last = rnorm;
      SGCALL(run_cycle(s, phi, rhs, l_max, sp, iter, 2));
      iter++;
      SGCALL(fetch_scalar(c, 2, &rnorm)); // the stop test needs the norm on the host: one sync per V-cycle
      if (hist) hist[iter] = rnorm;
      goNorm = rnorm > sp->norm_thresh; goRedu = rnorm > sp->eps * initial; goIter = iter < sp->max_iter;
      goHang = iter < sp->imin || rnorm < (1 - sp->hang) * norm_last; goMin = iter < sp->iter_min;
    }
    exit_status = int(!goRedu) + int(!goIter) * 2 + int(!goHang) * 4 + int(!goNorm) * 8;
    CK(cudaEventRecord(e1, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  CK(cudaGetLastError());
  if (stats) {
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double per = 0;
    SGCALL(sg_solver_cell_updates_per_cycle(s, sp, &per));
    stats->iterations = iter; stats->exit_status = exit_status;
    stats->initial_resnorm = initial; stats->final_resnorm = rnorm;
    stats->cell_updates = per * iter; stats->device_ms = ms;
    stats->kernel_launches = c->launches - l0;
  }
  return SG_OK;
}

#include "sg_linear_host.inc"
