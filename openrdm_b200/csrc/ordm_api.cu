// ordm_api.cu -- the C ABI of include/openrdm_b200.h on top of the sm_100a kernels.
//
// Host orchestration of OpenRDM's MC-PDFT energy path (energy.cc:27-188):
//   K1 k1_density  ->  K2 k2_ontop  ->  K3 k3_translate_xc  ->  K4 k4_finalize  (-> NCCL)
// all on one CUDA stream per context.  There is no CPU implementation of any stage here:
// without a CUDA device every compute entry point returns ORDM_ERR_CUDA.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/openrdm_b200.h"
#include "../../include/ordm_synth.h"
#include "k0_ao2mo.cuh"
#include "k1_density.cuh"
#include "k1_tile.cuh"
#include "k2_ontop.cuh"
#include "k2_int8.cuh"
#include "k2_int8_pair.cuh"
#include "k3_xc.cuh"
#include "rawio.h"
#include "rdm_pack.h"
#include "reduce.cuh"
#include "synth_rdm.h"

namespace {

thread_local std::string g_last_error;

constexpr int RETRY_FP64 = -1000;   // internal: the int8 on-top kernel's error guard tripped, repeat the call with the FP64 kernels

constexpr size_t PT_ALIGN = 64;   // shard boundaries (ordm_shard_range) are multiples of 64 points
// Device leading dimensions, the weight / per-point vectors and the staging batches are multiples of the widest
// point tile any kernel moves with one bulk copy (k2_ontop<8,4>: 8 * 4 * 8 = 256 points; the CTA-pair int8 kernel:
// 2 * 128), so that the last tile of every column stays inside the allocation (padding is zero-filled).
constexpr size_t LD_ALIGN = 256;
inline size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct GridView {  // one resident grid or one streamed batch
  size_t npts = 0, ld = 0;
  int n = 0;
  bool gga = false;
  const double *w = nullptr, *phi = nullptr, *phix = nullptr, *phiy = nullptr, *phiz = nullptr;
  double* vec[ORDM_VEC_COUNT] = {};  // per-point outputs (null = not materialised)
  double *gx = nullptr, *gy = nullptr, *gz = nullptr, *pi_core = nullptr;
  double* pcg[3] = {nullptr, nullptr, nullptr};  // closed-form core part of grad Pi
  double* core[4] = {nullptr, nullptr, nullptr, nullptr};  // k1_core partials rc, cx, cy, cz
};

// ---- NCCL through dlopen (no link-time dependency; absent NCCL only breaks comm_attach) ----
struct IdBlob { char b[128]; };  // ncclUniqueId (nccl.h:37-38), passed by value
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, IdBlob, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

bool load_nccl(std::string& why) {
  if (g_nccl.handle) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle) break;
  }
  if (!g_nccl.handle) { why = std::string("dlopen libnccl.so.2 failed: ") + dlerror(); return false; }
  g_nccl.GetUniqueId = (int (*)(void*))dlsym(g_nccl.handle, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void**, int, IdBlob, int))dlsym(g_nccl.handle, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(g_nccl.handle, "ncclAllReduce");
  g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(g_nccl.handle, "ncclAllGather");
  g_nccl.CommDestroy = (int (*)(void*))dlsym(g_nccl.handle, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.handle, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
    why = "libnccl is missing a required symbol";
    return false;
  }
  return true;
}

}  // namespace

struct ordm_ctx {
  uint64_t ctx_id = 0;   // unique per process (never reused, unlike the address)
  int device = 0;
  int sm_count = 148;
  int max_smem_optin = 0;
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  cudaStream_t side_stream = nullptr;  // K1 runs here underneath K2 (ordm_energy, overlapped mode)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_k1[2] = {nullptr, nullptr};
  bool serial_stages = false;          // ORDM_SERIAL_STAGES=1: K1 -> K2 -> K3 back to back on one stream
  bool no_split = false;               // ORDM_NO_SPLIT=1: the whole K1 underneath K2 (lab / test hook)
  bool own_stream = true;
  unsigned opts = 0;
  double omega = 0.0;  // MCPDFT_OMEGA
  std::string err;

  // resident grid
  GridView g;
  double *d_w = nullptr, *d_phi[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t cap_pts = 0;  // allocated ld
  int cap_n = 0;
  bool cap_gga = false;
  size_t vec_cap = 0;
  bool vec_diag = false;

  // RDMs
  int ncore = 0, nact = 0;
  bool have_rdm1 = false, have_rdm2 = false, core_form = false;
  uint64_t rdm1_version = 0;  // bumped by every install_rdm1 (K1's constant-bank upload is cached on it)
  std::vector<double> h_Dsa, h_Dsb;  // symmetrised, row-major stride k1::MAX_ACT (or nact when generic)
  std::vector<double> h_A1a, h_A1b;  // the active 1-RDM blocks as given (nact x nact column-major): polyradical analysis
  double *d_Dsa_g = nullptr, *d_Dsb_g = nullptr;
  // K1t (k1_tile.cuh): Dsa + Dsb | Dsa - Dsb, stride k1::MAX_ACT, uploaded on the launching stream when the version moves
  std::vector<double> h_dmat;
  double* d_dmat = nullptr;
  uint64_t dmat_version = 0;
  bool have_diff = false;
  uint32_t k1t_in_call = 0;   // k1_tile launches since the last result was delivered (ordm_result.density_tiled)
  bool k1_tile = true;        // ORDM_K1_TILE=0: the thread-per-point k1_density for every request (lab / test hook)
  struct { const double* phi[4] = {nullptr, nullptr, nullptr, nullptr}; size_t ld = 0, npts = 0; int c0 = -1, ncols = 0; k1t::Maps maps; } tmap;   // K1t's tensor maps are re-encoded when the view changes
  // K2 operands
  int k2_npg = 0, k2_mt = 0;
  k2::Params k2p{};
  double* d_dfrag = nullptr;
  void* d_tasks = nullptr;
  size_t cap_dfrag = 0, cap_tasks = 0, cap_pairs = 0, cap_timg = 0, cap_timg_pair = 0;   // bytes allocated (ensure_dev)
  bool k2_ws = false, k2_force_plain = false;
  // K2 on the int8 tensor cores (k2_int8.cuh): 18..20 active orbitals, Pi only (grad Pi keeps the DMMA kernel)
  int k2_i8_mode = 1;          // ORDM_K2_INT8: 0 off, 1 on with 26 slice pairs (default), 2 on with 21 slice pairs
  bool k2_i8 = false, i8_no_overlap = false;   // ORDM_I8_NO_OVERLAP=1: stages back to back with the int8 K2 (lab / test hook)
  // k2_i8 = k2_i8_cfg (the RDMs' size and the build allow the int8 kernel) && !i8_blocked (its error guard flagged
  // points of the current inputs: the FP64 kernels take over until the RDMs or the orbitals change)
  bool k2_i8_cfg = false, i8_blocked = false;
  uint32_t guard_points = 0;   // points the guard flagged in the call that switched to the FP64 kernels
  k2i::Params k2ip{};
  uint8_t* d_timg = nullptr;
  // CTA-pair form of the int8 kernel (k2_int8_pair.cuh; ORDM_K2_PAIR=0 keeps the single-CTA kernel)
  bool k2_pair = true;
  k2i::Params k2pp{};
  uint8_t* d_timg_pair = nullptr;
  size_t k2p_smem = 0;
  int* d_i8_fail = nullptr;
  size_t k2i_smem = 0;
  int core_keep = -1, core_keep_pct = 6;   // core orbitals left to the density kernel in split mode (ORDM_CORE_KEEP, ORDM_CORE_KEEP_PCT)
  bool core_quads = true;     // ORDM_CORE_QUADS=0: two points per thread (k1_core_pairs) instead of four (k1_core_quads)
  bool core_pairs = true;     // ORDM_CORE_PAIRS=0: the deep core kernel with 8-byte loads (k1_core_light<., 8>)
  bool res_by_copy = false;   // ORDM_RES_COPY=1: result scalars by cudaMemcpyAsync instead of stores from the final kernel (lab)
  bool big_overlap = true;    // ORDM_BIG_OVERLAP=0 switches it off: the deep core kernel also beside a DMMA K2 that fills shared memory (26 active orbitals)
  bool k1act_under = false;   // ORDM_K1ACT_UNDER=1: K1's active-space part also runs beside the int8 K2 (lab hook)
  bool k1_split = false;  // two-kernel K1 (k1_core + active) measured slower than the fused kernel; kept as a test hook
  uint16_t* d_pairs = nullptr;
  // grad-Pi variant of K2: the whole symmetric Dt in fragment order (built on first use)
  k2::Params k2gp{};
  double* d_dfrag_full = nullptr;
  void* d_tasks_full = nullptr;
  int k2g_nbuf = 0;
  size_t k2g_smem = 0;
  bool k2g_ready = false;
  uint64_t k2_dmma_per_mtile = 0;
  size_t k2_smem = 0;
  std::vector<double> h_Dt;  // folded Dt kept for repacking when the tile shape changes

  // reductions
  double *d_partial = nullptr, *d_res = nullptr, *h_res = nullptr;
  double* d_partial3 = nullptr;   // K3's block sums in the resident path (K1's stay in d_partial: one final kernel reads both)
  size_t partial_cap = 0;
  // a step enqueued by ordm_energy_async and not yet collected by ordm_energy_wait
  struct Pending { bool active = false, overlap = false, split = false, ft = false; int functional = 0, launches = 0; double eclass = 0.0; } pend;
  cudaEvent_t ev[8] = {};

  // streamed (host-buffer) path
  struct Stage {
    double *w = nullptr, *phi[4] = {nullptr, nullptr, nullptr, nullptr};
    double* vec[16] = {};  // rho pi gx gy gz pi_core + k1_core partials (4) + grad Pi (3) + its core part (3)
    cudaEvent_t copied = nullptr, done = nullptr;
  } stage[2];
  size_t stage_pts = 0;
  int stage_n = 0;
  bool stage_gga = false;

  // stage bookkeeping
  bool rho_done = false, pi_done = false, xc_done = false, xc_ft = false, pig_done = false;
  // vectors installed by the caller through ordm_set_vector (MCPDFT::set_R, set_rhoa_x, ...: mcpdft.h:125-153)
  bool inj_R = false;      // translate / fully_translate read the caller's R instead of build_R's
  bool inj_grad = false;   // a spin-density gradient component changed: total gradient re-summed before the next translate
  double last_n[3] = {0, 0, 0}, last_trn[3] = {0, 0, 0};

  // NCCL
  void* comm = nullptr;
  int nranks = 1, rank = 0;
  // peer-memory all-reduce of the result scalars (reduce.cuh, k_allreduce_peer); falls back to NCCL
  bool peer_ok = false;
  ordm::PeerXchg px{};
  double* d_xchg = nullptr;
  int* d_xerr = nullptr;
  unsigned long long xepoch = 0;
};

namespace {

int fail(ordm_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  if (ctx) ctx->err = buf;
  return code;
}

#define CU(ctx, call)                                                                          \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      return fail(ctx, ORDM_ERR_CUDA, "CUDA error '%s' in %s (%s:%d)", cudaGetErrorString(e_), \
                  #call, __FILE__, __LINE__);                                                  \
  } while (0)

#define CHECK_CTX(ctx) \
  if (!(ctx)) return fail(nullptr, ORDM_ERR_INVALID, "null context")

template <typename T>
void dfree(T*& p) {
  if (p) cudaFree(p);
  p = nullptr;
}

// Device buffer that is re-used across RDM updates (an SCF loop calls ordm_set_active_space every iteration: a
// cudaFree / cudaMalloc pair per operand costs more than packing and uploading it)
template <typename T>
cudaError_t ensure_dev(T*& p, size_t& cap, size_t bytes) {
  if (p && bytes <= cap) return cudaSuccess;
  if (p) cudaFree(p);
  p = nullptr; cap = 0;
  const cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
  if (e == cudaSuccess) cap = bytes;
  return e;
}

// grad Pi is built when asked for, or when the fully-translated GGA gradients need it
bool want_pig(const ordm_ctx* c, bool gga) {
  return gga && (c->opts & (ORDM_OPT_PI_GRADIENT | ORDM_OPT_FULLY_TRANSLATED)) != 0;
}

int ensure_partials(ordm_ctx* c, size_t doubles) {
  if (doubles <= c->partial_cap) return ORDM_OK;
  dfree(c->d_partial);
  dfree(c->d_partial3);
  CU(c, cudaMalloc(&c->d_partial, doubles * sizeof(double)));
  CU(c, cudaMalloc(&c->d_partial3, doubles * sizeof(double)));
  c->partial_cap = doubles;
  return ORDM_OK;
}

// (re)allocate the resident per-point vectors according to the option flags
int ensure_vectors(ordm_ctx* c) {
  const bool diag = (c->opts & ORDM_OPT_DIAGNOSTICS) != 0;
  const bool pig = want_pig(c, c->g.gga);
  GridView& g = c->g;
  if (c->vec_cap >= g.ld && c->vec_diag == diag && g.vec[ORDM_VEC_RHO] && (!pig || g.vec[ORDM_VEC_PI_X])) return ORDM_OK;
  for (int i = 0; i < ORDM_VEC_COUNT; ++i)
    if (i != ORDM_VEC_W) dfree(g.vec[i]);
  dfree(g.gx); dfree(g.gy); dfree(g.gz); dfree(g.pi_core);
  for (auto& q : g.core) dfree(q);
  for (auto& q : g.pcg) dfree(q);
  const size_t bytes = g.ld * sizeof(double);
  auto al = [&](double*& p) -> cudaError_t { cudaError_t e = cudaMalloc(&p, bytes); if (e == cudaSuccess) e = cudaMemsetAsync(p, 0, bytes, c->stream); return e; };
  CU(c, al(g.vec[ORDM_VEC_RHO]));
  CU(c, al(g.vec[ORDM_VEC_PI]));
  CU(c, al(g.gx)); CU(c, al(g.gy)); CU(c, al(g.gz));
  CU(c, al(g.pi_core));
  for (auto& q : g.core) CU(c, al(q));
  if (diag) {
    for (int i = 0; i < ORDM_VEC_COUNT; ++i) {
      if (i == ORDM_VEC_W || i == ORDM_VEC_RHO || i == ORDM_VEC_PI) continue;
      if (i >= ORDM_VEC_PI_X && i <= ORDM_VEC_PI_Z && !pig) continue;
      CU(c, al(g.vec[i]));
    }
  } else if (pig) {
    CU(c, al(g.vec[ORDM_VEC_PI_X])); CU(c, al(g.vec[ORDM_VEC_PI_Y])); CU(c, al(g.vec[ORDM_VEC_PI_Z]));
  }
  if (pig)
    for (auto& q : g.pcg) CU(c, al(q));
  c->vec_cap = g.ld;
  c->vec_diag = diag;
  return ORDM_OK;
}

// ------------------------------------------------------------------ kernel launchers
template <bool GGA>
int launch_k1_np(ordm_ctx* c, const k1::Params& p, int grid, cudaStream_t st, bool under_k2 = false) {
  const int np = (int)round_up((size_t)std::max(c->nact, 1), 4);
  switch (np) {
    // under_k2: the kernel is to run beside the int8 K2, which holds the SM at the maximum shared-memory carve-out;
    // an SM does not re-partition while a CTA is resident, so this kernel has to ask for the same carve-out
#define K1CASE(N) case N: \
    if (under_k2) CU(c, cudaFuncSetAttribute(k1::k1_density<N, GGA>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); \
    k1::k1_density<N, GGA><<<grid, k1::THREADS, 0, st>>>(p); break;
    K1CASE(4) K1CASE(8) K1CASE(12) K1CASE(16) K1CASE(20) K1CASE(24) K1CASE(28) K1CASE(32)
#undef K1CASE
    default: k1::k1_density_generic<GGA><<<grid, k1::THREADS, 0, st>>>(p); break;
  }
  CU(c, cudaGetLastError());
  return ORDM_OK;
}

// ---- K1t: the active-space part on TMA-staged tiles (k1_tile.cuh).  Returns K1T_NO when the request is not one it
// serves (diagnostic vectors, grad Pi's core part, too many staged columns, unaligned views, no tensor-map entry point):
// the caller then launches k1_density.
constexpr int K1T_NO = -2000;
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TmapEncodeFn tmap_encoder() {
  static TmapEncodeFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return (TmapEncodeFn)f;
  }();
  return fn;
}

template <int NP>
int launch_k1_tile(ordm_ctx* c, const k1t::Maps& maps, const k1t::Params& p, int grid, size_t smem, cudaStream_t st) {
  static thread_local int dev = -1;
  static thread_local size_t bytes = 0;
  if (dev != c->device || bytes < smem) {
    CU(c, cudaFuncSetAttribute(k1t::k1_tile<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dev = c->device; bytes = smem;
  }
  k1t::k1_tile<NP><<<grid, k1t::THREADS, smem, st>>>(maps, p);
  CU(c, cudaGetLastError());
  return ORDM_OK;
}

int run_k1_tile(ordm_ctx* c, const GridView& v, const k1::Params& kp, cudaStream_t st, int* grid_out) {
  // (without orbital gradients a tile is one 1 KB-row chunk of ~20 columns and the thread-per-point kernel streams as fast:
  //  0.249 vs 0.267 ms per 2 M points, profiles/r02_k1_tile_lab.log; the tiled kernel serves the GGA layouts)
  if (!c->k1_tile || !v.gga || c->nact < 1 || c->nact > 28 || c->d_dmat == nullptr) return K1T_NO;
  if (kp.rhoa || kp.rhob || kp.saa || kp.sab || kp.sbb || kp.pcg[0] || !kp.rho || (v.gga && !kp.gx)) return K1T_NO;
  for (int i = 0; i < 6; ++i) if (kp.grad[i]) return K1T_NO;
  const int c0 = kp.core_rc ? (kp.core_done > 0 ? kp.core_done : c->ncore) : 0;   // first staged column
  const int nlead = c->ncore - c0, ncols = nlead + c->nact, np = (int)round_up((size_t)c->nact, 4);
  if (nlead < 0 || nlead > k1t::MAX_LEAD) return K1T_NO;
  const size_t fixed = k1t::dmat_doubles(np) * sizeof(double) + 16 * 12 + 1024 /* static shared memory */ + 1024 /* reserved per CTA */;
  const size_t slot = k1t::slot_doubles(ncols) * sizeof(double);
  if ((size_t)c->max_smem_optin < fixed + slot) return K1T_NO;
  const int nslots = (int)std::min<size_t>(12, ((size_t)c->max_smem_optin - fixed) / slot);
  if (nslots < 5) return K1T_NO;   // a tile's four chunks are held together + one in flight
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  if (!al16(v.phi) || (v.gga && (!al16(v.phix) || !al16(v.phiy) || !al16(v.phiz))) || !al16(v.w) || v.ld % 2 != 0 ||
      (kp.core_rc && (!al16(kp.core_rc) || !al16(kp.core_cx) || !al16(kp.core_cy) || !al16(kp.core_cz))) ||
      v.npts > (size_t)0x7fffff00u)
    return K1T_NO;
  TmapEncodeFn enc = tmap_encoder();
  if (!enc) return K1T_NO;
  auto& tm = c->tmap;
  const double* base[4] = {v.phi, v.phix, v.phiy, v.phiz};
  if (tm.phi[0] != base[0] || tm.phi[1] != base[1] || tm.phi[2] != base[2] || tm.phi[3] != base[3] || tm.ld != v.ld || tm.npts != v.npts ||
      tm.c0 != c0 || tm.ncols != ncols) {
    for (int a = 0; a < 4; ++a) {
      cuuint64_t gdim[2] = {(cuuint64_t)v.npts, (cuuint64_t)ncols}, gstr[1] = {(cuuint64_t)v.ld * 8};
      cuuint32_t box[2] = {(cuuint32_t)k1t::TILE, (cuuint32_t)ncols}, estr[2] = {1, 1};
      const CUresult r = enc(&tm.maps.m[a], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base[a]) + (size_t)c0 * v.ld, gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { tm.phi[0] = nullptr; return K1T_NO; }
    }
    for (int a = 0; a < 4; ++a) tm.phi[a] = base[a];
    tm.ld = v.ld; tm.npts = v.npts; tm.c0 = c0; tm.ncols = ncols;
  }
  if (c->dmat_version != c->rdm1_version) {
    CU(c, cudaMemcpyAsync(c->d_dmat, c->h_dmat.data(), c->h_dmat.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    c->dmat_version = c->rdm1_version;
  }
  k1t::Params p{};
  p.w = v.w; p.core_rc = kp.core_rc; p.core_cx = kp.core_cx; p.core_cy = kp.core_cy; p.core_cz = kp.core_cz;
  p.npts = v.npts; p.ntiles = (int)((v.npts + k1t::TILE - 1) / k1t::TILE);
  p.nlead = nlead; p.nact = c->nact; p.have_diff = c->have_diff ? 1 : 0; p.nslots = nslots;
  p.rho = kp.rho; p.gx = kp.gx; p.gy = kp.gy; p.gz = kp.gz; p.pi_core = kp.pi_core;
  p.partial = kp.partial; p.dmat = c->d_dmat;
  const int grid = std::max(1, std::min(p.ntiles, c->sm_count));
  const size_t smem = k1t::smem_bytes(ncols, nslots, np);
  int rc = ORDM_OK;
  switch (np) {
#define K1TCASE(N) case N: rc = launch_k1_tile<N>(c, tm.maps, p, grid, smem, st); break;
    K1TCASE(4) K1TCASE(8) K1TCASE(12) K1TCASE(16) K1TCASE(20) K1TCASE(24) K1TCASE(28)
#undef K1TCASE
    default: return K1T_NO;
  }
  if (rc) return rc;
  *grid_out = grid;
  c->k1t_in_call += 1;
  return ORDM_OK;
}

int k1_grid(const ordm_ctx* c, size_t npts) {
  const size_t want = (npts + k1::THREADS - 1) / k1::THREADS;
  return (int)std::max<size_t>(1, std::min<size_t>(want, (size_t)c->sm_count * 16));
}
int k3_grid(const ordm_ctx* c, size_t npts) {
  const size_t want = (npts + k3::THREADS - 1) / k3::THREADS;
  return (int)std::max<size_t>(1, std::min<size_t>(want, (size_t)c->sm_count * 8));
}

// The core sums alone (rc, cx, cy, cz -> v.core[]), in the register-light shape that runs underneath K2.
// How many of the core orbitals the side kernel sums while K2 runs; the rest is left to the density kernel that follows.
// Beside the int8 K2 the side kernel's DFMAs only get the FP64 pipe between the MMAs (the pipe stalls 13x while
// tcgen05.mma runs), so it ends after K2 does: handing the last few columns to the density kernel,
// which has the pipe to itself, shortens the step.  ORDM_CORE_KEEP=<n> overrides the number left behind.
int core_done_beside_k2(const ordm_ctx* c) {
  int keep = c->core_keep >= 0 ? c->core_keep : (c->k2_i8 ? (c->ncore * c->core_keep_pct + 50) / 100 : 0);
  keep = std::min(keep, std::max(c->ncore - 8, 0));
  return c->ncore - keep;   // (= ncore: the side kernel does them all)
}

int run_k1_core_light(ordm_ctx* c, const GridView& v, int* launches, cudaStream_t st, bool deep = false) {
  k1::Params p{};
  p.phi = v.phi; p.phix = v.phix; p.phiy = v.phiy; p.phiz = v.phiz;
  p.ld = v.ld; p.npts = v.npts; p.ncore = c->ncore; p.nact = c->nact;
  p.core_rc = v.core[0]; p.core_cx = v.core[1]; p.core_cy = v.core[2]; p.core_cz = v.core[3];
  p.core_done = core_done_beside_k2(c);
  const size_t want = (v.npts + k1::LIGHT_THREADS - 1) / k1::LIGHT_THREADS;
  const int grid = (int)std::max<size_t>(1, std::min<size_t>(want, (size_t)c->sm_count * (deep ? 1 : 16)));
  if (deep) {   // next to the int8 K2: one CTA per SM with 32 loads in flight per thread
    // K2 holds the SM at the maximum shared-memory carve-out, and an SM does not re-partition while a CTA is
    // resident: ask for the same carve-out or this kernel waits for K2 to leave
    static thread_local int carve_dev = -1;
    if (carve_dev != c->device) {
      CU(c, cudaFuncSetAttribute(k1::k1_core_light<true, 8>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      CU(c, cudaFuncSetAttribute(k1::k1_core_light<false, 8>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      CU(c, cudaFuncSetAttribute(k1::k1_core_pairs<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      CU(c, cudaFuncSetAttribute(k1::k1_core_pairs<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      carve_dev = c->device;
    }
    auto al32 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31u) == 0; };
    if (c->core_quads && v.ld % 4 == 0 && al32(v.phi) && al32(v.phix) && al32(v.phiy) && al32(v.phiz) && al32(v.core[0]) && al32(v.core[1]) && al32(v.core[2]) &&
        al32(v.core[3])) {   // four points per thread, 32-byte loads
      static thread_local int carve_q = -1;
      if (carve_q != c->device) {
        CU(c, cudaFuncSetAttribute(k1::k1_core_quads<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CU(c, cudaFuncSetAttribute(k1::k1_core_quads<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        carve_q = c->device;
      }
      if (v.gga) k1::k1_core_quads<true><<<grid, k1::LIGHT_THREADS, 0, st>>>(p);
      else k1::k1_core_quads<false><<<grid, k1::LIGHT_THREADS, 0, st>>>(p);
      CU(c, cudaGetLastError());
      *launches += 1;
      return ORDM_OK;
    }
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    if (c->core_pairs && v.ld % 2 == 0 && al16(v.phi) && al16(v.phix) && al16(v.phiy) && al16(v.phiz) && al16(v.core[0])) {   // two points per thread, 16-byte loads (the arrays are ld long and 16-byte aligned)
      if (v.gga) k1::k1_core_pairs<true><<<grid, k1::LIGHT_THREADS, 0, st>>>(p);
      else k1::k1_core_pairs<false><<<grid, k1::LIGHT_THREADS, 0, st>>>(p);
    } else if (v.gga) k1::k1_core_light<true, 8><<<grid, k1::LIGHT_THREADS, 0, st>>>(p);
    else k1::k1_core_light<false, 8><<<grid, k1::LIGHT_THREADS, 0, st>>>(p);
  } else if (v.gga) k1::k1_core_light<true, 2><<<grid, k1::LIGHT_THREADS, 0, st>>>(p);
  else k1::k1_core_light<false, 2><<<grid, k1::LIGHT_THREADS, 0, st>>>(p);
  CU(c, cudaGetLastError());
  *launches += 1;
  return ORDM_OK;
}

// core_ready: the core sums are already in v.core[] (run_k1_core_light); only the active-space part and
// the assembly of rho, grad rho, pi_core are left
// final_grid: when given, the one-block final pass is left to the caller (k4_final_all) and *final_grid receives the
// number of partial rows in d_partial
int run_k1(ordm_ctx* c, const GridView& v, int accumulate, int* launches, cudaStream_t st = nullptr, bool core_ready = false, bool under_k2 = false,
           int* final_grid = nullptr) {
  if (!st) st = c->stream;
  // The __constant__ bank holding Ds is per device and shared by every context of the process
  // (kernel-parameter space was tried instead: ptxas then LDCs the coefficients into registers,
  // 254 regs, K1 2.6 -> 3.4 ms).  The bank is uploaded only when another context, another RDM or
  // another stream used it last.  The lock is held from the ownership check to the K1 launch, and
  // the last user's K1 is followed by an event that the next uploader's stream waits for: two
  // contexts on one device (two host threads) serialise their K1s on the GPU instead of racing.
  k1::Params p{};
  p.phi = v.phi; p.phix = v.phix; p.phiy = v.phiy; p.phiz = v.phiz; p.w = v.w;
  p.ld = v.ld; p.npts = v.npts; p.ncore = c->ncore; p.nact = c->nact;
  p.rho = v.vec[ORDM_VEC_RHO];
  p.gx = v.gx; p.gy = v.gy; p.gz = v.gz;
  p.pi_core = c->core_form ? v.pi_core : nullptr;
  for (int i = 0; i < 3; ++i) p.pcg[i] = (c->core_form && v.gga && want_pig(c, v.gga)) ? v.pcg[i] : nullptr;
  p.rhoa = v.vec[ORDM_VEC_RHOA]; p.rhob = v.vec[ORDM_VEC_RHOB];
  for (int i = 0; i < 6; ++i) p.grad[i] = v.vec[ORDM_VEC_RHOA_X + i];
  p.saa = v.vec[ORDM_VEC_SIGMA_AA]; p.sab = v.vec[ORDM_VEC_SIGMA_AB]; p.sbb = v.vec[ORDM_VEC_SIGMA_BB];
  p.Dsa_g = c->d_Dsa_g; p.Dsb_g = c->d_Dsb_g;
  if (core_ready) {
    p.core_rc = v.core[0]; p.core_cx = v.core[1]; p.core_cy = v.core[2]; p.core_cz = v.core[3];
    p.core_done = core_done_beside_k2(c);
    } else if (c->k1_split && c->ncore >= 8 && v.core[0]) {  // K1a: stream the core columns at full occupancy first
    p.core_rc = v.core[0]; p.core_cx = v.core[1]; p.core_cy = v.core[2]; p.core_cz = v.core[3];
    const size_t want = (v.npts + k1::CORE_THREADS - 1) / k1::CORE_THREADS;
    const int cgrid = (int)std::max<size_t>(1, std::min<size_t>(want, (size_t)c->sm_count * 8));
    if (v.gga) k1::k1_core<true><<<cgrid, k1::CORE_THREADS, 0, st>>>(p);
    else k1::k1_core<false><<<cgrid, k1::CORE_THREADS, 0, st>>>(p);
    CU(c, cudaGetLastError());
    *launches += 1;
  }
  int rc = ensure_partials(c, (size_t)c->sm_count * 16 * 8);
  if (rc) return rc;
  p.partial = c->d_partial;
  if (!under_k2 && st != c->side_stream) {   // the tiled kernel takes the energy-path requests; it needs the SM's shared memory, so not beside K2 (side stream)
    int tgrid = 0;
    rc = run_k1_tile(c, v, p, st, &tgrid);
    if (rc == ORDM_OK) {
      if (final_grid) { *final_grid = tgrid; *launches += 1; return ORDM_OK; }
      ordm::k4_finalize<3><<<1, 256, 0, st>>>(c->d_partial, tgrid, c->d_res, ordm::RES_N, accumulate);
      CU(c, cudaGetLastError());
      *launches += 2;
      return ORDM_OK;
    }
    if (rc != K1T_NO) return rc;
  }
  static std::mutex bank_mu;
  static struct { uint64_t owner_id; uint64_t version; cudaStream_t stream; cudaEvent_t last_k1; } bank[64] = {};
  std::unique_lock<std::mutex> bank_lock(bank_mu, std::defer_lock);
  const bool use_bank = c->nact <= k1::MAX_ACT;
  if (use_bank) {
    bank_lock.lock();
    auto& b = bank[c->device & 63];
    if (!b.last_k1) CU(c, cudaEventCreateWithFlags(&b.last_k1, cudaEventDisableTiming));
    if (b.owner_id != c->ctx_id || b.version != c->rdm1_version || b.stream != st) {
      if (b.owner_id && b.stream != st) CU(c, cudaStreamWaitEvent(st, b.last_k1, 0));  // the last user's K1 may still read the bank
      CU(c, cudaMemcpyToSymbolAsync(k1::c_Dsa, c->h_Dsa.data(), sizeof(double) * k1::MAX_ACT * k1::MAX_ACT, 0, cudaMemcpyHostToDevice, st));
      CU(c, cudaMemcpyToSymbolAsync(k1::c_Dsb, c->h_Dsb.data(), sizeof(double) * k1::MAX_ACT * k1::MAX_ACT, 0, cudaMemcpyHostToDevice, st));
      b.owner_id = c->ctx_id; b.version = c->rdm1_version; b.stream = st;
    }
  }
  const int grid = k1_grid(c, v.npts);
  rc = v.gga ? launch_k1_np<true>(c, p, grid, st, under_k2) : launch_k1_np<false>(c, p, grid, st, under_k2);
  if (rc) return rc;
  if (use_bank) {
    CU(c, cudaEventRecord(bank[c->device & 63].last_k1, st));
    bank_lock.unlock();
  }
  if (final_grid) { *final_grid = grid; *launches += 1; return ORDM_OK; }
  ordm::k4_finalize<3><<<1, 256, 0, st>>>(c->d_partial, grid, c->d_res, ordm::RES_N, accumulate);
  CU(c, cudaGetLastError());
  *launches += 2;
  return ORDM_OK;
}

template <int NPG, int MT>
int launch_k2(ordm_ctx* c, k2::Params p) {
  constexpr int P = 8 * MT * NPG;
  p.ntiles = (int)((p.npts + P - 1) / P);
  const int grid = std::max(1, std::min(p.ntiles, c->sm_count));
  static thread_local int attr_dev = -1;
  static thread_local size_t attr_bytes = 0;
  if (attr_dev != c->device || attr_bytes < c->k2_smem) {
    CU(c, cudaFuncSetAttribute(k2::k2_ontop<NPG, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->k2_smem));
    attr_dev = c->device;
    attr_bytes = c->k2_smem;
  }
  k2::k2_ontop<NPG, MT><<<grid, k2::THREADS, c->k2_smem, c->stream>>>(p);
  CU(c, cudaGetLastError());
  return ORDM_OK;
}

template <int NBUF>
int launch_k2_grad(ordm_ctx* c, const k2::Params& p, int grid) {
  static thread_local int dev = -1;
  static thread_local size_t bytes = 0;
  if (dev != c->device || bytes < c->k2g_smem) {
    CU(c, cudaFuncSetAttribute(k2::k2_ontop_ws<true, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->k2g_smem));
    dev = c->device; bytes = c->k2g_smem;
  }
  k2::k2_ontop_ws<true, NBUF><<<grid, k2::WS_THREADS, c->k2g_smem, c->stream>>>(p);
  CU(c, cudaGetLastError());
  return ORDM_OK;
}

int configure_k2_grad(ordm_ctx* c);

// Pi and grad Pi in one pass (full Dt): MCPDFT::build_pi for a GGA case, mcpdft.cc:178-183
int run_k2_grad(ordm_ctx* c, const GridView& v, int* launches) {
  int rc = configure_k2_grad(c);
  if (rc) return rc;
  k2::Params p = c->k2gp;
  p.phi_act = v.phi + (size_t)c->ncore * v.ld;
  p.dphi_act[0] = v.phix + (size_t)c->ncore * v.ld;
  p.dphi_act[1] = v.phiy + (size_t)c->ncore * v.ld;
  p.dphi_act[2] = v.phiz + (size_t)c->ncore * v.ld;
  p.ld = v.ld;
  p.npts = v.npts;
  p.pi_core = c->core_form ? v.pi_core : nullptr;
  p.pi = v.vec[ORDM_VEC_PI];
  for (int i = 0; i < 3; ++i) {
    p.pi_core_g[i] = c->core_form ? v.pcg[i] : nullptr;
    p.pi_g[i] = v.vec[ORDM_VEC_PI_X + i];
  }
  p.ntiles = (int)((p.npts + k2::WS_P - 1) / k2::WS_P);
  const int grid = std::max(1, std::min(p.ntiles, c->sm_count));
  rc = c->k2g_nbuf == 2 ? launch_k2_grad<2>(c, p, grid) : launch_k2_grad<1>(c, p, grid);
  if (rc) return rc;
  *launches += 1;
  return ORDM_OK;
}

template <int NA, int GMIN>
int launch_k2_int8(ordm_ctx* c, const k2i::Params& p, int grid) {
  static thread_local int dev = -1;
  static thread_local size_t bytes = 0;
  if (dev != c->device || bytes < c->k2i_smem) {
    CU(c, cudaFuncSetAttribute(k2i::k2i_ontop<NA, GMIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->k2i_smem));
    dev = c->device; bytes = c->k2i_smem;
  }
  k2i::k2i_ontop<NA, GMIN><<<grid, k2i::THREADS, c->k2i_smem, c->stream>>>(p);
  CU(c, cudaGetLastError());
  return ORDM_OK;
}

constexpr int K2P_EW = 2;   // epilogue warps per TMEM lane quadrant of the CTA-pair kernel (9 warps, allocated as 12, x 128 registers
                            // leave 16 K registers to the co-resident k1_core_light CTA)
template <int NA, int GMIN>
int launch_k2_pair(ordm_ctx* c, const k2i::Params& p, int grid) {
  static thread_local int dev = -1;
  static thread_local size_t bytes = 0;
  if (dev != c->device || bytes < c->k2p_smem) {
    CU(c, cudaFuncSetAttribute(k2i::k2i_ontop_pair<NA, GMIN, K2P_EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->k2p_smem));
    dev = c->device; bytes = c->k2p_smem;
  }
  k2i::k2i_ontop_pair<NA, GMIN, K2P_EW><<<grid, (4 * K2P_EW + 1) * 32, c->k2p_smem, c->stream>>>(p);   // clusters of 2 (__cluster_dims__)
  CU(c, cudaGetLastError());
  return ORDM_OK;
}

int run_k2_int8_pair(ordm_ctx* c, const GridView& v, int* launches, bool lean) {
  k2i::Params p = c->k2pp;
  p.phi_act = v.phi + (size_t)c->ncore * v.ld;
  p.ld = v.ld;
  p.npts = v.npts;
  p.pi_core = (c->core_form && !lean) ? v.pi_core : nullptr;
  p.pi = v.vec[ORDM_VEC_PI];
  const size_t pair_tiles = (v.npts + 2 * k2i::TILE - 1) / (2 * k2i::TILE);
  const int grid = 2 * (int)std::max<size_t>(1, std::min<size_t>(pair_tiles, (size_t)c->sm_count / 2));
  const bool wide = c->k2_i8_mode == 1;   // 26 slice pairs
  int rc;
  switch (c->nact) {
    case 18: rc = wide ? launch_k2_pair<18, 4>(c, p, grid) : launch_k2_pair<18, 5>(c, p, grid); break;
    case 19: rc = wide ? launch_k2_pair<19, 4>(c, p, grid) : launch_k2_pair<19, 5>(c, p, grid); break;
    case 20: rc = wide ? launch_k2_pair<20, 4>(c, p, grid) : launch_k2_pair<20, 5>(c, p, grid); break;
    default: return fail(c, ORDM_ERR_UNSUPPORTED, "int8 on-top path selected for %d active orbitals", c->nact);
  }
  if (rc) return rc;
  *launches += 1;
  return ORDM_OK;
}

// Pi of the active space on the int8 tensor cores (18..20 active orbitals); lean as in run_k2: no pi_core add
int run_k2_int8(ordm_ctx* c, const GridView& v, int* launches, bool lean) {
  if (c->k2_pair) return run_k2_int8_pair(c, v, launches, lean);
  k2i::Params p = c->k2ip;
  p.phi_act = v.phi + (size_t)c->ncore * v.ld;
  p.ld = v.ld;
  p.npts = v.npts;
  p.pi_core = (c->core_form && !lean) ? v.pi_core : nullptr;
  p.pi = v.vec[ORDM_VEC_PI];
  const int grid = (int)std::max<size_t>(1, std::min<size_t>((v.npts + k2i::TILE - 1) / k2i::TILE, (size_t)c->sm_count));
  const bool wide = c->k2_i8_mode == 1;   // 26 slice pairs
  int rc;
  switch (c->nact) {
    case 18: rc = wide ? launch_k2_int8<18, 4>(c, p, grid) : launch_k2_int8<18, 5>(c, p, grid); break;
    case 19: rc = wide ? launch_k2_int8<19, 4>(c, p, grid) : launch_k2_int8<19, 5>(c, p, grid); break;
    case 20: rc = wide ? launch_k2_int8<20, 4>(c, p, grid) : launch_k2_int8<20, 5>(c, p, grid); break;
    default: return fail(c, ORDM_ERR_UNSUPPORTED, "int8 on-top path selected for %d active orbitals", c->nact);
  }
  if (rc) return rc;
  *launches += 1;
  return ORDM_OK;
}

// lean = the register-trimmed WS kernel without the pi_core add (K1 is still running underneath;
// K3 completes Pi)
int run_k2(ordm_ctx* c, const GridView& v, int* launches, bool lean = false) {
  if (want_pig(c, v.gga)) return run_k2_grad(c, v, launches);
  if (c->k2_i8) return run_k2_int8(c, v, launches, lean);
  k2::Params p = c->k2p;
  p.phi_act = v.phi + (size_t)c->ncore * v.ld;
  p.ld = v.ld;
  p.npts = v.npts;
  p.pi_core = (c->core_form && !lean) ? v.pi_core : nullptr;
  p.pi = v.vec[ORDM_VEC_PI];
  int rc;
  if (c->k2_ws && lean) {
    p.ntiles = (int)((p.npts + k2::WS_P - 1) / k2::WS_P);
    const int grid = std::max(1, std::min(p.ntiles, c->sm_count));
    static thread_local int ln_dev = -1;
    static thread_local size_t ln_bytes = 0;
    if (ln_dev != c->device || ln_bytes < c->k2_smem) {
      CU(c, cudaFuncSetAttribute(k2::k2_ontop_ws<false, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->k2_smem));
      ln_dev = c->device; ln_bytes = c->k2_smem;
    }
    k2::k2_ontop_ws<false, 2, true><<<grid, k2::WS_THREADS, c->k2_smem, c->stream>>>(p);
    CU(c, cudaGetLastError());
    rc = ORDM_OK;
  } else if (c->k2_ws) {
    p.ntiles = (int)((p.npts + k2::WS_P - 1) / k2::WS_P);
    const int grid = std::max(1, std::min(p.ntiles, c->sm_count));
    static thread_local int ws_dev = -1;
    static thread_local size_t ws_bytes = 0;
    if (ws_dev != c->device || ws_bytes < c->k2_smem) {
      CU(c, cudaFuncSetAttribute(k2::k2_ontop_ws<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->k2_smem));
      ws_dev = c->device; ws_bytes = c->k2_smem;
    }
    k2::k2_ontop_ws<false, 2><<<grid, k2::WS_THREADS, c->k2_smem, c->stream>>>(p);
    CU(c, cudaGetLastError());
    rc = ORDM_OK;
  } else if (c->k2_npg == 8 && c->k2_mt == 4) rc = launch_k2<8, 4>(c, p);
  else if (c->k2_npg == 4 && c->k2_mt == 4) rc = launch_k2<4, 4>(c, p);
  else if (c->k2_npg == 2 && c->k2_mt == 4) rc = launch_k2<2, 4>(c, p);
  else if (c->k2_npg == 1 && c->k2_mt == 4) rc = launch_k2<1, 4>(c, p);
  else if (c->k2_npg == 1 && c->k2_mt == 2) rc = launch_k2<1, 2>(c, p);
  else rc = launch_k2<1, 1>(c, p);
  if (rc) return rc;
  *launches += 1;
  return ORDM_OK;
}

int run_k3(ordm_ctx* c, const GridView& v, int functional, bool ft, int accumulate, int* launches, bool add_core = false,
           const double* R_in = nullptr, int* final_grid = nullptr) {
  k3::Params p{};
  p.rho = v.vec[ORDM_VEC_RHO]; p.pi = v.vec[ORDM_VEC_PI]; p.w = v.w;
  p.R_in = R_in;
  if (add_core && c->core_form) { p.pi_core = v.pi_core; p.pi_out = v.vec[ORDM_VEC_PI]; }
  p.omega = c->omega;
  p.gx = v.gga ? v.gx : nullptr; p.gy = v.gga ? v.gy : nullptr; p.gz = v.gga ? v.gz : nullptr;
  p.npts = v.npts;
  p.R = v.vec[ORDM_VEC_R]; p.tra = v.vec[ORDM_VEC_TR_RHOA]; p.trb = v.vec[ORDM_VEC_TR_RHOB];
  p.tsaa = v.vec[ORDM_VEC_TR_SIGMA_AA]; p.tsab = v.vec[ORDM_VEC_TR_SIGMA_AB]; p.tsbb = v.vec[ORDM_VEC_TR_SIGMA_BB];
  const int grid = k3_grid(c, v.npts);
  int rc = ensure_partials(c, (size_t)c->sm_count * 16 * 8);
  if (rc) return rc;
  p.partial = final_grid ? c->d_partial3 : c->d_partial;
  // energy.cc:114-120: "SVWN" -> LSDA+VWN3, else PBE.  Without orbital gradients the
  // translated sigma are zero vectors (energy.cc:63-65): PBE then runs with gx=gy=gz=0.
  if (functional != ORDM_FUNC_SVWN) {
    if (!v.gga) { p.gx = p.gy = p.gz = nullptr; }
    if (ft && v.gga) { p.px = v.vec[ORDM_VEC_PI_X]; p.py = v.vec[ORDM_VEC_PI_Y]; p.pz = v.vec[ORDM_VEC_PI_Z]; }
  }
#define K3CASE(F)                                                                           \
  case F:                                                                                   \
    if (ft) k3::k3_translate_xc<F, true><<<grid, k3::THREADS, 0, c->stream>>>(p);            \
    else k3::k3_translate_xc<F, false><<<grid, k3::THREADS, 0, c->stream>>>(p);              \
    break;
  switch (functional) {
    K3CASE(0) K3CASE(1) K3CASE(2) K3CASE(3) K3CASE(4) K3CASE(5)
    default: return fail(c, ORDM_ERR_INVALID, "unknown functional id %d", functional);
  }
#undef K3CASE
  CU(c, cudaGetLastError());
  if (final_grid) { *final_grid = grid; *launches += 1; return ORDM_OK; }
  ordm::k4_finalize<5><<<1, 256, 0, c->stream>>>(c->d_partial, grid, c->d_res, ordm::RES_TRN, accumulate);
  CU(c, cudaGetLastError());
  *launches += 2;
  return ORDM_OK;
}

// choose the K2 tile shape that fits shared memory and (re)pack Dt' for its lane count
int configure_k2(ordm_ctx* c) {
  const int M = ordm::pair_count(c->nact);
  const int nblk = (M + k2::BLK - 1) / k2::BLK, Mpad = nblk * k2::BLK;
  struct Cand { int npg, mt; size_t bytes; } cand[4] = {
      {2, 4, k2::Smem<64>::bytes(Mpad, c->nact)}, {1, 4, k2::Smem<32>::bytes(Mpad, c->nact)},
      {1, 2, k2::Smem<16>::bytes(Mpad, c->nact)}, {1, 1, k2::Smem<8>::bytes(Mpad, c->nact)}};
  int pick = -1;
  for (int i = 0; i < 4; ++i)
    if (cand[i].bytes <= (size_t)c->max_smem_optin) { pick = i; break; }
  c->k2_ws = !c->k2_force_plain && k2::WsSmem::bytes(Mpad, c->nact) <= (size_t)c->max_smem_optin;
  if (c->k2_ws) { cand[1].bytes = k2::WsSmem::bytes(Mpad, c->nact); pick = 1; }  // 32-point tiles, 8 task lanes
  // Small active spaces: Dt has too few 32x32 blocks to feed eight task lanes (M = 36 -> 3 blocks), so
  // the warps split POINTS instead: 8 (or 4) point groups of 32 points per tile, 1 (or 2) task lanes.
  const int nblocks = nblk * (nblk + 1) / 2;
  Cand wide8{8, 4, k2::Smem<256>::bytes(Mpad, c->nact)}, wide4{4, 4, k2::Smem<128>::bytes(Mpad, c->nact)};
  if (!c->k2_force_plain && nblocks <= 3 && wide8.bytes <= (size_t)c->max_smem_optin) { cand[0] = wide8; pick = 0; c->k2_ws = false; }
  else if (!c->k2_force_plain && nblocks <= 6 && wide4.bytes <= (size_t)c->max_smem_optin) { cand[0] = wide4; pick = 0; c->k2_ws = false; }
  if (pick < 0)
    return fail(c, ORDM_ERR_UNSUPPORTED, "on-top pair density: %d orbitals (M=%d pairs) exceed shared memory; use ordm_set_active_space with a smaller active block", c->nact, M);
  c->k2_npg = cand[pick].npg;
  c->k2_mt = cand[pick].mt;
  c->k2_smem = cand[pick].bytes;
  ordm::DfragLayout lay;
  ordm::pack_dfrag(M, c->h_Dt, k2::WARPS / c->k2_npg, k2::PREFETCH, lay);
  static_assert(sizeof(ordm::SegDesc) == sizeof(k2::SegDesc) && sizeof(ordm::SegDesc) == 8, "segment descriptor layout");
  std::vector<uint16_t> tab;
  ordm::pair_table(c->nact, tab);
  // the operands are overwritten in place: nothing enqueued earlier may still be reading them
  CU(c, cudaStreamSynchronize(c->stream));
  if (c->side_stream) CU(c, cudaStreamSynchronize(c->side_stream));
  CU(c, ensure_dev(c->d_pairs, c->cap_pairs, tab.size() * sizeof(uint16_t)));
  CU(c, cudaMemcpy(c->d_pairs, tab.data(), tab.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
  CU(c, ensure_dev(c->d_dfrag, c->cap_dfrag, lay.data.size() * sizeof(double)));
  CU(c, ensure_dev(c->d_tasks, c->cap_tasks, lay.segs.size() * sizeof(ordm::SegDesc)));
  CU(c, cudaMemcpy(c->d_dfrag, lay.data.data(), lay.data.size() * sizeof(double), cudaMemcpyHostToDevice));
  CU(c, cudaMemcpy(c->d_tasks, lay.segs.data(), lay.segs.size() * sizeof(ordm::SegDesc), cudaMemcpyHostToDevice));
  c->k2p = k2::Params{};
  c->k2p.nact = c->nact; c->k2p.M = M; c->k2p.nblk = nblk;
  c->k2p.pairs = c->d_pairs;
  c->k2p.dfrag = c->d_dfrag; c->k2p.segs = reinterpret_cast<const k2::SegDesc*>(c->d_tasks);
  for (int i = 0; i <= k2::MAX_LANES; ++i) { c->k2p.seg_ofs[i] = lay.seg_ofs[i]; c->k2p.blk_ofs[i] = lay.blk_ofs[i]; }
  c->k2_dmma_per_mtile = lay.subtiles;
  c->k2g_ready = false;  // the grad-Pi packing follows the same Dt: rebuilt on next use
  // the int8 tensor-core path for the active-space sizes where K2 dominates the step
  const int kp = (M + 31) / 32 * 32;
  c->k2_i8_cfg = c->k2_i8_mode > 0 && c->nact >= k2i::MIN_ACT && c->nact <= k2i::MAX_ACT &&
                 k2i::Smem::bytes(kp, c->nact) <= (size_t)c->max_smem_optin;
  c->i8_blocked = false;
  c->k2_i8 = c->k2_i8_cfg;
  if (c->k2_i8) {
    static_assert(k2i::NSLICE == 6 && k2i::FIX_BITS == 46, "ordm::pack_t_int8 is written for six byte slices of a 46-bit image");
    int t_exp = 0;
    c->k2_pair = c->k2_pair && k2i::PairSmem::bytes(kp) + 1024 <= (size_t)c->max_smem_optin;
    if (!c->d_i8_fail) { CU(c, cudaMalloc(&c->d_i8_fail, 2 * sizeof(int))); CU(c, cudaMemset(c->d_i8_fail, 0, 2 * sizeof(int))); }
    c->k2ip = k2i::Params{};
    if (c->k2_pair) {   // the image cut in two: each CTA of a pair keeps half of the rows of every chunk (the single-CTA image is not built)
      std::vector<uint8_t> pimg;
      ordm::pack_t_int8_pair(M, c->h_Dt, kp, pimg, t_exp);
      if (pimg.size() != 2 * k2i::PairSmem::t_bytes(kp))
        return fail(c, ORDM_ERR_UNSUPPORTED, "int8 on-top path: pair image and kernel geometry disagree");
      CU(c, ensure_dev(c->d_timg_pair, c->cap_timg_pair, pimg.size()));
      CU(c, cudaMemcpy(c->d_timg_pair, pimg.data(), pimg.size(), cudaMemcpyHostToDevice));
      c->k2p_smem = k2i::PairSmem::bytes(kp);
    } else {
      std::vector<uint8_t> img;
      ordm::pack_t_int8(M, c->h_Dt, kp, img, t_exp);
      if (img.size() != k2i::Smem::t_bytes(kp) || ordm::t_int8_slice_bytes(kp) != k2i::Geom::slice_bytes(kp))
        return fail(c, ORDM_ERR_UNSUPPORTED, "int8 on-top path: host image and kernel geometry disagree");
      CU(c, ensure_dev(c->d_timg, c->cap_timg, img.size()));
      CU(c, cudaMemcpy(c->d_timg, img.data(), img.size(), cudaMemcpyHostToDevice));
    }
    c->k2ip.timg = c->d_timg; c->k2ip.t_exp = t_exp; c->k2ip.fail = c->d_i8_fail;
    ordm::int8_guard_constants(M, c->h_Dt, t_exp, c->k2_i8_mode == 1 ? 4 : 5, c->k2ip.guard_g1, c->k2ip.guard_g2);
    c->k2pp = c->k2ip;
    c->k2pp.timg = c->d_timg_pair;
    c->k2i_smem = k2i::Smem::bytes(kp, c->nact);
  }
  return ORDM_OK;
}

int configure_k2_grad(ordm_ctx* c) {
  if (c->k2g_ready) return ORDM_OK;
  const int M = ordm::pair_count(c->nact);
  const int nblk = (M + k2::BLK - 1) / k2::BLK, Mpad = nblk * k2::BLK;
  c->k2g_nbuf = 0;
  for (int nbuf = 2; nbuf >= 1; --nbuf)
    if (k2::WsSmem::bytes(Mpad, c->nact, true, nbuf) <= (size_t)c->max_smem_optin) { c->k2g_nbuf = nbuf; break; }
  if (!c->k2g_nbuf)
    return fail(c, ORDM_ERR_UNSUPPORTED, "grad Pi: %d active orbitals (M=%d pairs) exceed shared memory", c->nact, M);
  c->k2g_smem = k2::WsSmem::bytes(Mpad, c->nact, true, c->k2g_nbuf);
  ordm::DfragLayout lay;
  ordm::pack_dfrag(M, c->h_Dt, k2::WARPS, k2::PREFETCH, lay, /*full=*/true);
  dfree(c->d_dfrag_full);
  dfree(c->d_tasks_full);
  CU(c, cudaMalloc(&c->d_dfrag_full, lay.data.size() * sizeof(double)));
  CU(c, cudaMalloc(&c->d_tasks_full, lay.segs.size() * sizeof(ordm::SegDesc)));
  CU(c, cudaMemcpy(c->d_dfrag_full, lay.data.data(), lay.data.size() * sizeof(double), cudaMemcpyHostToDevice));
  CU(c, cudaMemcpy(c->d_tasks_full, lay.segs.data(), lay.segs.size() * sizeof(ordm::SegDesc), cudaMemcpyHostToDevice));
  c->k2gp = k2::Params{};
  c->k2gp.nact = c->nact; c->k2gp.M = M; c->k2gp.nblk = nblk;
  c->k2gp.pairs = c->d_pairs;
  c->k2gp.dfrag = c->d_dfrag_full; c->k2gp.segs = reinterpret_cast<const k2::SegDesc*>(c->d_tasks_full);
  for (int i = 0; i <= k2::MAX_LANES; ++i) { c->k2gp.seg_ofs[i] = lay.seg_ofs[i]; c->k2gp.blk_ofs[i] = lay.blk_ofs[i]; }
  c->k2g_ready = true;
  return ORDM_OK;
}

int install_rdm1(ordm_ctx* c, int ncore, int nact, const double* A1a, const double* A1b) {
  c->ncore = ncore;
  c->nact = nact;
  const int stride = (nact <= k1::MAX_ACT) ? k1::MAX_ACT : nact;
  ordm::symmetrise_opdm(nact, A1a, stride, c->h_Dsa);
  ordm::symmetrise_opdm(nact, A1b, stride, c->h_Dsb);
  c->h_A1a.assign(A1a, A1a + (size_t)nact * nact);
  c->h_A1b.assign(A1b, A1b + (size_t)nact * nact);
  dfree(c->d_Dsa_g);
  dfree(c->d_Dsb_g);
  if (nact > k1::MAX_ACT) {
    const size_t b = (size_t)nact * nact * sizeof(double);
    CU(c, cudaMalloc(&c->d_Dsa_g, b));
    CU(c, cudaMalloc(&c->d_Dsb_g, b));
    CU(c, cudaMemcpy(c->d_Dsa_g, c->h_Dsa.data(), b, cudaMemcpyHostToDevice));
    CU(c, cudaMemcpy(c->d_Dsb_g, c->h_Dsb.data(), b, cudaMemcpyHostToDevice));
  }
  if (nact <= k1::MAX_ACT) {   // K1t's coefficient block: Dsa + Dsb | Dsa - Dsb (both symmetric), stride k1::MAX_ACT
    const size_t m = (size_t)k1::MAX_ACT * k1::MAX_ACT;
    c->h_dmat.assign(2 * m, 0.0);
    c->have_diff = false;
    for (size_t i = 0; i < m; ++i) {
      c->h_dmat[i] = c->h_Dsa[i] + c->h_Dsb[i];
      c->h_dmat[m + i] = c->h_Dsa[i] - c->h_Dsb[i];
      if (c->h_dmat[m + i] != 0.0) c->have_diff = true;
    }
    if (!c->d_dmat) CU(c, cudaMalloc(&c->d_dmat, 2 * m * sizeof(double)));
  }
  c->have_rdm1 = true;
  static std::atomic<uint64_t> next_version{0};
  c->rdm1_version = ++next_version;
  c->rho_done = c->pi_done = c->xc_done = c->pig_done = false;
  return ORDM_OK;
}

int check_inputs(ordm_ctx* c, bool need_rdm2) {
  if (!c->g.w || !c->g.phi) return fail(c, ORDM_ERR_STATE, "grid/orbitals not set (ordm_set_grid, ordm_set_orbitals)");
  if (!c->have_rdm1) return fail(c, ORDM_ERR_STATE, "1-RDMs not set (ordm_set_rdm1 / ordm_set_active_space)");
  if (need_rdm2 && !c->have_rdm2) return fail(c, ORDM_ERR_STATE, "2-RDM not set (ordm_set_rdm2ab_dense / ordm_set_active_space)");
  if (c->ncore + c->nact != c->g.n)
    return fail(c, ORDM_ERR_STATE, "RDM dimension %d+%d does not match the %d orbital columns", c->ncore, c->nact, c->g.n);
  return ORDM_OK;
}

// the D2H read of the result scalars (and of the int8 kernel's / the peer all-reduce's status words), stream-ordered
int fetch_enqueue(ordm_ctx* c, bool written_by_kernel = false) {
  if (written_by_kernel) {   // k4_final_all stored scalars and status words into the pinned block itself
    CU(c, cudaEventRecord(c->ev[5], c->stream));
    return ORDM_OK;
  }
  CU(c, cudaMemcpyAsync(c->h_res, c->d_res, ordm::RES_COUNT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  int* h_i8_fail = reinterpret_cast<int*>(c->h_res + 12);   // the pinned block has room beyond the RES_COUNT scalars
  if (c->k2_i8 && c->d_i8_fail) CU(c, cudaMemcpyAsync(h_i8_fail, c->d_i8_fail, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  if (c->comm && c->peer_ok && c->d_xerr) CU(c, cudaMemcpyAsync(h_i8_fail + 2, c->d_xerr, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaEventRecord(c->ev[5], c->stream));
  return ORDM_OK;
}

int fetch_finish(ordm_ctx* c, double eclass, ordm_result* out, uint64_t npts, int launches) {
  int* h_i8_fail = reinterpret_cast<int*>(c->h_res + 12);
  int* h_xerr = h_i8_fail + 2;
  CU(c, cudaStreamSynchronize(c->stream));
  if (!(c->k2_i8 && c->d_i8_fail)) h_i8_fail[0] = h_i8_fail[1] = 0;
  if (!(c->comm && c->peer_ok && c->d_xerr)) *h_xerr = 0;
  if (*h_xerr)   // k_allreduce_peer (reduce.cuh): a peer's flag never arrived; the result scalars were set to NaN
    return fail(c, ORDM_ERR_NCCL, "peer all-reduce timed out waiting for another rank (rank %d of %d)", c->rank, c->nranks);
  if (*h_i8_fail) {   // a tcgen05 completion barrier never flipped (bounded spin in k2_int8.cuh): the Pi of this call is not valid
    cudaMemsetAsync(c->d_i8_fail, 0, 2 * sizeof(int), c->stream);
    return fail(c, ORDM_ERR_CUDA, "int8 on-top kernel: %d CTA(s) timed out waiting for a tensor-core completion barrier", *h_i8_fail);
  }
  if (h_i8_fail[1]) {   // the error guard flagged points: this call's Pi is not trusted; the caller repeats it with the FP64 kernels
    cudaMemsetAsync(c->d_i8_fail, 0, 2 * sizeof(int), c->stream);
    c->guard_points = (uint32_t)h_i8_fail[1];
    c->i8_blocked = true;
    c->k2_i8 = false;
    return RETRY_FP64;
  }
  const double* r = c->h_res;
  std::memset(out, 0, sizeof(*out));
  out->eclass = eclass;
  out->ex = r[ordm::RES_EX];
  out->ec = r[ordm::RES_EC];
  out->e_tot = 0.0; out->e_tot += eclass; out->e_tot += out->ex; out->e_tot += out->ec;  // energy.cc:181-183
  out->n_tot = r[ordm::RES_N]; out->n_a = r[ordm::RES_NA]; out->n_b = r[ordm::RES_NB];
  out->trn_tot = r[ordm::RES_TRN]; out->trn_a = r[ordm::RES_TRNA]; out->trn_b = r[ordm::RES_TRNB];
  out->npts = npts;
  out->gpu_launches = (uint32_t)launches;
  out->density_tiled = c->k1t_in_call;
  c->k1t_in_call = 0;
  for (int i = 0; i < 3; ++i) { c->last_n[i] = r[ordm::RES_N + i]; c->last_trn[i] = r[ordm::RES_TRN + i]; }
  return ORDM_OK;
}

int fetch_result(ordm_ctx* c, double eclass, ordm_result* out, uint64_t npts, int launches) {
  int rc = fetch_enqueue(c);
  return rc ? rc : fetch_finish(c, eclass, out, npts, launches);
}

// all-reduce of d_res over the attached communicator: NVLink peer stores when the exchange buffers
// could be mapped, NCCL otherwise
int all_reduce_results(ordm_ctx* c, int* launches) {
  if (!c->comm) return ORDM_OK;
  if (c->peer_ok) {
    ordm::k_allreduce_peer<<<1, 128, 0, c->stream>>>(c->d_res, c->px, ++c->xepoch, c->d_xerr);
    CU(c, cudaGetLastError());
  } else {
    int nr = g_nccl.AllReduce(c->d_res, c->d_res, ordm::RES_COUNT, /*ncclDouble*/ 8, /*ncclSum*/ 0, c->comm, c->stream);
    if (nr != 0) return fail(c, ORDM_ERR_NCCL, "ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(nr) : "?");
  }
  *launches += 1;
  return ORDM_OK;
}

void release_peer_xchg(ordm_ctx* c) {
  if (c->peer_ok)
    for (int r = 0; r < c->nranks; ++r)
      if (r != c->rank && c->px.peer[r]) cudaIpcCloseMemHandle(c->px.peer[r]);
  c->peer_ok = false;
  c->px = ordm::PeerXchg{};
  dfree(c->d_xchg);
  dfree(c->d_xerr);
}

// exchange CUDA-IPC handles of the per-rank exchange buffers over the fresh NCCL communicator and map
// every peer's buffer; a failure on any rank just leaves the NCCL path in place on all of them
int setup_peer_xchg(ordm_ctx* c) {
  release_peer_xchg(c);
  if (getenv("ORDM_NO_PEER_ALLREDUCE") || c->nranks < 2 || c->nranks > 16 || !g_nccl.AllGather) return ORDM_OK;
  const size_t bytes = ordm::peer_xchg_bytes(c->nranks);
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof(mine));
  // the two scratch buffers every rank needs to take part in the collectives below
  char* d_h = nullptr;
  int* d_flag = nullptr;
  // d_flag is the one buffer the gathers below cannot do without: if even that fails, this rank still has to issue
  // the same collectives as its peers (they would hang in ncclAllGather otherwise), so fall back to the result
  // vector as scratch -- 16 doubles hold nranks + 1 <= 17 ints -- and report "not ok"
  const bool scratch_ok = cudaMalloc(&d_h, sizeof(mine) * (size_t)(c->nranks + 1)) == cudaSuccess &&
                          cudaMalloc(&d_flag, sizeof(int) * (size_t)(c->nranks + 1)) == cudaSuccess;
  cudaGetLastError();
  int* flag_buf = d_flag ? d_flag : reinterpret_cast<int*>(c->d_res);
  int ok = scratch_ok && cudaMalloc(&c->d_xchg, bytes) == cudaSuccess && cudaMemset(c->d_xchg, 0, bytes) == cudaSuccess &&
           cudaMalloc(&c->d_xerr, sizeof(int)) == cudaSuccess && cudaMemset(c->d_xerr, 0, sizeof(int)) == cudaSuccess &&
           cudaIpcGetMemHandle(&mine, c->d_xchg) == cudaSuccess;
  cudaGetLastError();
  std::vector<cudaIpcMemHandle_t> all(c->nranks);
  std::vector<int> flags(c->nranks, 0);
  auto gather_flag = [&](int v) -> bool {  // all-gather one int per rank into flags[]
    if (cudaMemcpy(flag_buf, &v, sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) return false;
    if (g_nccl.AllGather(flag_buf, flag_buf + 1, sizeof(int), /*ncclChar*/ 0, c->comm, c->stream) != 0) return false;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return false;
    return cudaMemcpy(flags.data(), flag_buf + 1, sizeof(int) * c->nranks, cudaMemcpyDeviceToHost) == cudaSuccess;
  };
  auto all_set = [&]() { for (int r = 0; r < c->nranks; ++r) if (flags[r] != 1) return false; return true; };
  bool good = gather_flag(ok) && all_set();
  // the handles travel only if EVERY rank got its buffers (all ranks see the same flags, so all take the same branch)
  bool coll = false;
  if (good) {
    cudaMemcpy(d_h, &mine, sizeof(mine), cudaMemcpyHostToDevice);
    coll = g_nccl.AllGather(d_h, d_h + sizeof(mine), sizeof(mine), 0, c->comm, c->stream) == 0 && cudaStreamSynchronize(c->stream) == cudaSuccess &&
           cudaMemcpy(all.data(), d_h + sizeof(mine), sizeof(mine) * c->nranks, cudaMemcpyDeviceToHost) == cudaSuccess;
  }
  int mapped = 0;
  if (good && coll) {
    c->px.nranks = c->nranks; c->px.rank = c->rank;
    c->px.peer[c->rank] = c->d_xchg;
    mapped = 1;
    for (int r = 0; r < c->nranks; ++r) {
      if (r == c->rank) continue;
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); mapped = 0; break; }
      c->px.peer[r] = (double*)p;
    }
  }
  // the fast path is used only if EVERY rank mapped every peer
  const bool everyone = gather_flag(mapped) && all_set();
  const bool used_res_as_scratch = (d_flag == nullptr);
  dfree(d_h);
  dfree(d_flag);
  cudaGetLastError();
  if (used_res_as_scratch) cudaMemset(c->d_res, 0, 16 * sizeof(double));   // the result vector served as scratch
  if (everyone) { c->peer_ok = true; c->xepoch = 0; }
  else {
    for (int r = 0; r < c->nranks; ++r)
      if (r != c->rank && c->px.peer[r]) cudaIpcCloseMemHandle(c->px.peer[r]);
    c->px = ordm::PeerXchg{};
    release_peer_xchg(c);
  }
  return ORDM_OK;
}

// total density gradient from the spin components (MCPDFT::translate_density_gradients adds them per point,
// mcpdft.cc:383-388); only needed after the caller installed components through ordm_set_vector
__global__ void k_sum2(double* __restrict__ dst, const double* __restrict__ a, const double* __restrict__ b, size_t n) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) dst[p] = a[p] + b[p];
}

// row-major [m][ncols_src] (numpy / data.h5 layout of /PHI/PHI_MO) -> the first ncols columns, column-major with
// leading dimension ld: a 32 x 32 shared-memory transpose, coalesced on both sides
__global__ void k_rows_to_columns(double* __restrict__ dst, size_t ld, const double* __restrict__ src, size_t m, int ncols_src, int ncols) {
  __shared__ double tile[32][33];
  const size_t r0 = (size_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {   // read: consecutive threads along a source row
    const size_t r = r0 + i;
    const int cc = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < m && cc < ncols) ? src[r * (size_t)ncols_src + cc] : 0.0;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {   // write: consecutive threads along a destination column
    const int cc = c0 + i;
    const size_t r = r0 + threadIdx.x;
    if (r < m && cc < ncols) dst[(size_t)cc * ld + r] = tile[threadIdx.x][i];
  }
}

// synthetic fill kernels ----------------------------------------------------------------
__global__ void k_synth_phi(double* __restrict__ dst, size_t ld, size_t npts_local, size_t p0, size_t npts_total,
                            int n, uint64_t comp_key, uint64_t env_key) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npts_local) return;
  const double env = ordm_env_key(env_key, p0 + p);
  for (int i = blockIdx.y; i < n; i += gridDim.y) dst[(size_t)i * ld + p] = ordm_synth_phi(comp_key, env, npts_total, p0 + p, (uint64_t)i);
}
__global__ void k_synth_w(double* __restrict__ dst, size_t npts_local, size_t p0, double inv_total, uint64_t w_key) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < npts_local) dst[p] = ordm_synth_w(w_key, inv_total, p0 + p);
}

int alloc_grid(ordm_ctx* c, size_t npts, int n, bool gga) {
  const size_t ld = round_up(std::max<size_t>(npts, 1), LD_ALIGN);
  if (ld > c->cap_pts || n > c->cap_n || (gga && !c->cap_gga) || !c->d_phi[0]) {
    for (int k = 0; k < 4; ++k) dfree(c->d_phi[k]);
    const size_t bytes = ld * (size_t)n * sizeof(double);
    for (int k = 0; k < (gga ? 4 : 1); ++k) {
      CU(c, cudaMalloc(&c->d_phi[k], bytes));
      CU(c, cudaMemsetAsync(c->d_phi[k], 0, bytes, c->stream));
    }
    c->cap_pts = ld; c->cap_n = n; c->cap_gga = gga;
  }
  c->g.ld = c->cap_pts;
  c->g.n = n;
  c->g.gga = gga;
  c->g.npts = npts;
  c->g.phi = c->d_phi[0];
  c->g.phix = gga ? c->d_phi[1] : nullptr;
  c->g.phiy = gga ? c->d_phi[2] : nullptr;
  c->g.phiz = gga ? c->d_phi[3] : nullptr;
  c->rho_done = c->pi_done = c->xc_done = c->pig_done = false;
  c->i8_blocked = false; c->k2_i8 = c->k2_i8_cfg;   // the int8 guard is re-evaluated on the new orbitals
  return ORDM_OK;
}

int alloc_w(ordm_ctx* c, size_t npts) {
  const size_t ld = round_up(std::max<size_t>(npts, 1), LD_ALIGN);
  dfree(c->d_w);
  CU(c, cudaMalloc(&c->d_w, ld * sizeof(double)));
  CU(c, cudaMemsetAsync(c->d_w, 0, ld * sizeof(double), c->stream));
  c->g.w = c->d_w;
  c->g.vec[ORDM_VEC_W] = c->d_w;
  c->rho_done = c->pi_done = c->xc_done = c->pig_done = false;
  return ORDM_OK;
}

// K0 launcher: C[m x N] = A[m x K] B[K x N] on the context's stream (k0_ao2mo.cuh)
template <int NT>
int launch_k0(ordm_ctx* c, const k0::Params& p) {
  static thread_local int dev = -1;
  if (dev != c->device) {
    CU(c, cudaFuncSetAttribute(k0::k0_ao2mo<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k0::Smem<NT>::bytes));
    dev = c->device;
  }
  dim3 grid((unsigned)((p.m + k0::BM - 1) / k0::BM), (unsigned)((p.N + 8 * NT - 1) / (8 * NT)));
  k0::k0_ao2mo<NT><<<grid, k0::threads_of<NT>(), k0::Smem<NT>::bytes, c->stream>>>(p);
  CU(c, cudaGetLastError());
  return ORDM_OK;
}
int run_k0(ordm_ctx* c, const k0::Params& p) {
  if (p.m == 0) return ORDM_OK;
  if (p.N <= 32) return launch_k0<4>(c, p);
  if (p.N <= 64) return launch_k0<8>(c, p);
  return launch_k0<13>(c, p);   // 104 columns per CTA: config 4's 103 occupied MOs in one pass over A
}

// The AO values are streamed through the device in point batches (H2D on the copy stream overlapped with the GEMM of
// the previous batch).  row_major = 0: ao(p, mu) at ao[p + mu * ld] (the C ABI's convention); 1: at ao[p * ld + mu]
// (numpy / data.h5: /PHI/PHI_AO is [npts][nao]); C is nao x nmo column-major in both cases.
int ao_to_mo(ordm_ctx* c, size_t npts, int nao, size_t ld, const double* const src[4], bool gga, int nmo, const double* C, bool row_major) {
  int rc = alloc_grid(c, npts, nmo, gga);
  if (rc) return rc;
  if (!npts) return ORDM_OK;
  // point batches: two staging buffers of batch x nao doubles (<= 256 MB each)
  size_t batch = std::max<size_t>(LD_ALIGN, ((size_t)1 << 25) / (size_t)nao / LD_ALIGN * LD_ALIGN);
  batch = std::min(batch, round_up(npts, LD_ALIGN));
  double *d_C = nullptr, *d_stage[2] = {nullptr, nullptr};
  cudaEvent_t copied[2] = {nullptr, nullptr}, used[2] = {nullptr, nullptr};
  auto cleanup = [&]() {
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamSynchronize(c->stream);
    dfree(d_C);
    for (int b = 0; b < 2; ++b) { dfree(d_stage[b]); if (copied[b]) cudaEventDestroy(copied[b]); if (used[b]) cudaEventDestroy(used[b]); }
  };
#define CUX(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail(c, ORDM_ERR_CUDA, "CUDA error '%s' in %s", cudaGetErrorString(e_), #call); } } while (0)
  CUX(cudaMalloc(&d_C, (size_t)nao * nmo * sizeof(double)));
  CUX(cudaMemcpyAsync(d_C, C, (size_t)nao * nmo * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  for (int b = 0; b < 2; ++b) {
    CUX(cudaMalloc(&d_stage[b], batch * (size_t)nao * sizeof(double)));
    CUX(cudaEventCreateWithFlags(&copied[b], cudaEventDisableTiming));
    CUX(cudaEventCreateWithFlags(&used[b], cudaEventDisableTiming));
  }
  size_t it = 0;
  for (int k = 0; k < (gga ? 4 : 1); ++k)
    for (size_t p0 = 0; p0 < npts; p0 += batch, ++it) {
      const int b = (int)(it & 1);
      const size_t m = std::min(batch, npts - p0);
      if (it >= 2) CUX(cudaStreamWaitEvent(c->copy_stream, used[b], 0));
      if (row_major)   // m rows of nao values (row stride ld) -> packed [m][nao]
        CUX(cudaMemcpy2DAsync(d_stage[b], (size_t)nao * sizeof(double), src[k] + p0 * ld, ld * sizeof(double), (size_t)nao * sizeof(double), m,
                              cudaMemcpyHostToDevice, c->copy_stream));
      else
        CUX(cudaMemcpy2DAsync(d_stage[b], batch * sizeof(double), src[k] + p0, ld * sizeof(double), m * sizeof(double), (size_t)nao,
                              cudaMemcpyHostToDevice, c->copy_stream));
      CUX(cudaEventRecord(copied[b], c->copy_stream));
      CUX(cudaStreamWaitEvent(c->stream, copied[b], 0));
      // phi[p0:p0+m, 0:nmo] = stage[0:m, 0:nao] . C
      k0::Params gp{};
      gp.A = d_stage[b]; gp.lda = row_major ? (size_t)nao : batch; gp.a_row_major = row_major ? 1 : 0;
      gp.B = d_C; gp.C = c->d_phi[k] + p0; gp.ldc = c->g.ld; gp.m = m; gp.K = nao; gp.N = nmo;
      if ((rc = run_k0(c, gp))) { cleanup(); return rc; }
      CUX(cudaEventRecord(used[b], c->stream));
    }
  CUX(cudaStreamSynchronize(c->stream));
#undef CUX
  cleanup();
  return ORDM_OK;
}

}  // namespace

// =========================================================================================
//                                      C ABI
// =========================================================================================
extern "C" {

const char* ordm_version(void) { return "openrdm_b200 0.1 (sm_100a)"; }

int ordm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char* ordm_last_error(const ordm_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int ordm_create(int device, ordm_ctx** out) {
  if (!out) return fail(nullptr, ORDM_ERR_INVALID, "null output pointer");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, ORDM_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(nullptr, ORDM_ERR_INVALID, "device %d out of range [0,%d)", device, ndev);
  CU(nullptr, cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(nullptr, cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(nullptr, ORDM_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
  ordm_ctx* c = new ordm_ctx();
  static std::atomic<uint64_t> next_ctx_id{0};
  c->ctx_id = ++next_ctx_id;
  c->device = device;
  if (const char* e = getenv("ORDM_I8_NO_OVERLAP")) c->i8_no_overlap = (e[0] == '1');
  if (const char* e = getenv("ORDM_K2_INT8")) c->k2_i8_mode = atoi(e);   // 0: DMMA kernels only; 2: 21 slice pairs
  if (const char* e = getenv("ORDM_K1ACT_UNDER")) c->k1act_under = (e[0] == '1');
  if (const char* e = getenv("ORDM_K1_TILE")) c->k1_tile = atoi(e) != 0;
  if (const char* e = getenv("ORDM_CORE_QUADS")) c->core_quads = atoi(e) != 0;
  if (const char* e = getenv("ORDM_CORE_KEEP")) c->core_keep = atoi(e);
  if (const char* e = getenv("ORDM_CORE_KEEP_PCT")) c->core_keep_pct = atoi(e);
  if (const char* e = getenv("ORDM_CORE_PAIRS")) c->core_pairs = (e[0] == '1');
  if (const char* e = getenv("ORDM_RES_COPY")) c->res_by_copy = (e[0] == '1');
  if (const char* e = getenv("ORDM_BIG_OVERLAP")) c->big_overlap = (e[0] == '1');
  if (const char* e = getenv("ORDM_K2_PAIR")) c->k2_pair = (e[0] != '0');   // 0: single-CTA int8 kernel (k2_int8.cuh)
  if (const char* e = getenv("ORDM_K2_PLAIN")) c->k2_force_plain = (e[0] == '1');  // test hook: single-buffer K2
  if (const char* e = getenv("ORDM_K1_SPLIT")) c->k1_split = (e[0] == '1');         // test hook: two-kernel K1
  c->sm_count = prop.multiProcessorCount;
  c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  CU(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CU(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  CU(c, cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking));
  CU(c, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  CU(c, cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  for (auto& ev : c->ev_k1) CU(c, cudaEventCreate(&ev));
  if (const char* e = getenv("ORDM_SERIAL_STAGES")) c->serial_stages = (e[0] == '1');  // per-kernel timing / test hook
  if (const char* e = getenv("ORDM_NO_SPLIT")) c->no_split = (e[0] == '1');
  for (auto& ev : c->ev) CU(c, cudaEventCreate(&ev));
  CU(c, cudaMalloc(&c->d_res, 16 * sizeof(double)));
  CU(c, cudaMemset(c->d_res, 0, 16 * sizeof(double)));
  CU(c, cudaMallocHost(&c->h_res, 16 * sizeof(double)));
  *out = c;
  return ORDM_OK;
}

void ordm_destroy(ordm_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  release_peer_xchg(c);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  for (int i = 0; i < ORDM_VEC_COUNT; ++i)
    if (i != ORDM_VEC_W) dfree(c->g.vec[i]);
  dfree(c->g.gx); dfree(c->g.gy); dfree(c->g.gz); dfree(c->g.pi_core);
  for (auto& q : c->g.core) dfree(q);
  for (auto& q : c->g.pcg) dfree(q);
  dfree(c->d_dfrag_full); dfree(c->d_tasks_full);
  dfree(c->d_w);
  for (int k = 0; k < 4; ++k) dfree(c->d_phi[k]);
  dfree(c->d_dmat);
  dfree(c->d_Dsa_g); dfree(c->d_Dsb_g); dfree(c->d_dfrag); dfree(c->d_tasks); dfree(c->d_pairs);
  dfree(c->d_timg); dfree(c->d_timg_pair); dfree(c->d_i8_fail);
  dfree(c->d_partial); dfree(c->d_partial3); dfree(c->d_res);
  if (c->h_res) cudaFreeHost(c->h_res);
  for (auto& s : c->stage) {
    dfree(s.w);
    for (auto& p : s.phi) dfree(p);
    for (auto& p : s.vec) dfree(p);
    if (s.copied) cudaEventDestroy(s.copied);
    if (s.done) cudaEventDestroy(s.done);
  }
  for (auto& ev : c->ev)
    if (ev) cudaEventDestroy(ev);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->side_stream) cudaStreamDestroy(c->side_stream);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  for (auto& ev : c->ev_k1)
    if (ev) cudaEventDestroy(ev);
  delete c;
}

int ordm_set_stream(ordm_ctx* c, void* s) {
  CHECK_CTX(c);
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaStreamSynchronize(c->stream));
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  c->stream = (cudaStream_t)s;
  c->own_stream = false;
  return ORDM_OK;
}

int ordm_set_options(ordm_ctx* c, unsigned flags) {
  CHECK_CTX(c);
  if ((flags ^ c->opts) & ~(unsigned)ORDM_OPT_SERIAL_STAGES) c->rho_done = c->pi_done = c->xc_done = c->pig_done = false;
  c->opts = flags;
  return ORDM_OK;
}

int ordm_set_omega(ordm_ctx* c, double omega) {
  CHECK_CTX(c);
  c->omega = omega;
  return ORDM_OK;
}

int ordm_synchronize(ordm_ctx* c) {
  CHECK_CTX(c);
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaStreamSynchronize(c->stream));
  return ORDM_OK;
}

int ordm_host_alloc(size_t bytes, void** out) {
  if (!out) return fail(nullptr, ORDM_ERR_INVALID, "null output pointer");
  // ORDM_HOST_WC=1 (lab switch): write-combined pages -- not snooped during the DMA reads of an H2D copy; slow to read
  // back on the host, so only for buffers the host merely fills
  static const bool wc = [] { const char* e = getenv("ORDM_HOST_WC"); return e && e[0] == '1'; }();
  CU(nullptr, cudaHostAlloc(out, bytes ? bytes : 1, wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
  return ORDM_OK;
}
int ordm_host_free(void* p) {
  if (p) CU(nullptr, cudaFreeHost(p));
  return ORDM_OK;
}

// ---- inputs ------------------------------------------------------------------------------
int ordm_set_grid(ordm_ctx* c, size_t npts, const double* w) {
  CHECK_CTX(c);
  if (npts && !w) return fail(c, ORDM_ERR_INVALID, "null weights");
  CU(c, cudaSetDevice(c->device));
  int rc = alloc_w(c, npts);
  if (rc) return rc;
  if (c->g.phi && c->g.npts != npts) { c->g.phi = nullptr; }  // orbitals must be re-set for a new grid size
  c->g.npts = npts;
  if (npts) CU(c, cudaMemcpyAsync(c->d_w, w, npts * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return ORDM_OK;
}

int ordm_set_orbitals(ordm_ctx* c, size_t npts, int n, size_t ld, const double* phi, const double* px, const double* py, const double* pz) {
  CHECK_CTX(c);
  if (n <= 0 || (npts && !phi) || ld < npts) return fail(c, ORDM_ERR_INVALID, "bad orbital matrix (n=%d, ld=%zu < npts=%zu or null)", n, ld, npts);
  const bool gga = px || py || pz;
  if (gga && !(px && py && pz)) return fail(c, ORDM_ERR_INVALID, "give all of phi_x, phi_y, phi_z or none");
  if (c->g.w && c->g.npts != npts) return fail(c, ORDM_ERR_INVALID, "orbitals have %zu points but the grid has %zu", npts, c->g.npts);
  CU(c, cudaSetDevice(c->device));
  int rc = alloc_grid(c, npts, n, gga);
  if (rc) return rc;
  const double* src[4] = {phi, px, py, pz};
  if (npts)
    for (int k = 0; k < (gga ? 4 : 1); ++k)
      CU(c, cudaMemcpy2DAsync(c->d_phi[k], c->g.ld * sizeof(double), src[k], ld * sizeof(double), npts * sizeof(double), (size_t)n, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return ORDM_OK;
}

int ordm_set_orbitals_ao(ordm_ctx* c, size_t npts, int nao, size_t ld, const double* ao, const double* ax, const double* ay,
                         const double* az, int nmo, const double* C) {
  CHECK_CTX(c);
  if (nao <= 0 || nmo <= 0 || !C || (npts && !ao) || ld < npts) return fail(c, ORDM_ERR_INVALID, "bad AO matrix / MO coefficients (nao=%d nmo=%d ld=%zu npts=%zu)", nao, nmo, ld, npts);
  const bool gga = ax || ay || az;
  if (gga && !(ax && ay && az)) return fail(c, ORDM_ERR_INVALID, "give all of ao_x, ao_y, ao_z or none");
  if (c->g.w && c->g.npts != npts) return fail(c, ORDM_ERR_INVALID, "orbitals have %zu points but the grid has %zu", npts, c->g.npts);
  CU(c, cudaSetDevice(c->device));
  const double* src[4] = {ao, ax, ay, az};
  return ao_to_mo(c, npts, nao, ld, src, gga, nmo, C, /*row_major=*/false);
}

int ordm_set_rdm1(ordm_ctx* c, int n, const double* D1a, const double* D1b) {
  CHECK_CTX(c);
  if (n <= 0 || !D1a || !D1b) return fail(c, ORDM_ERR_INVALID, "bad 1-RDM");
  CU(c, cudaSetDevice(c->device));
  if (c->core_form || c->nact != n) c->have_rdm2 = false;  // a dense D2 of another dimension is stale
  c->core_form = false;
  return install_rdm1(c, 0, n, D1a, D1b);
}

int ordm_set_rdm2ab_dense(ordm_ctx* c, int n, const double* D2ab) {
  CHECK_CTX(c);
  if (n <= 0 || !D2ab) return fail(c, ORDM_ERR_INVALID, "bad 2-RDM");
  if (c->have_rdm1 && (c->core_form || c->nact != n))
    return fail(c, ORDM_ERR_INVALID, "dense D2ab of %d orbitals does not match the 1-RDMs (%d+%d)", n, c->ncore, c->nact);
  CU(c, cudaSetDevice(c->device));
  if (!c->have_rdm1) { c->ncore = 0; c->nact = n; }
  ordm::fold_tpdm(n, D2ab, c->h_Dt);
  int rc = configure_k2(c);
  if (rc) return rc;
  c->have_rdm2 = true;
  c->pi_done = c->xc_done = c->pig_done = false;
  return ORDM_OK;
}

int ordm_set_active_space(ordm_ctx* c, int ncore, int nact, const double* A1a, const double* A1b, const double* D2act) {
  CHECK_CTX(c);
  if (ncore < 0 || nact <= 0 || !A1a || !A1b) return fail(c, ORDM_ERR_INVALID, "bad active space (ncore=%d nact=%d)", ncore, nact);
  CU(c, cudaSetDevice(c->device));
  c->core_form = ncore > 0;
  int rc = install_rdm1(c, ncore, nact, A1a, A1b);
  if (rc) return rc;
  if (D2act) {
    ordm::fold_tpdm(nact, D2act, c->h_Dt);
    rc = configure_k2(c);
    if (rc) return rc;
    c->have_rdm2 = true;
  } else {
    c->have_rdm2 = false;
  }
  return ORDM_OK;
}

int ordm_set_active_space_sparse(ordm_ctx* c, int ncore, int nact, size_t nnz1a, const int* i1a, const int* j1a, const double* v1a,
                                 size_t nnz1b, const int* i1b, const int* j1b, const double* v1b, size_t nnz2, const int* i2,
                                 const int* j2, const int* k2, const int* l2, const double* v2, int upper_packed) {
  CHECK_CTX(c);
  if (ncore < 0 || nact <= 0) return fail(c, ORDM_ERR_INVALID, "bad active space (ncore=%d nact=%d)", ncore, nact);
  if ((nnz1a && !(i1a && j1a && v1a)) || (nnz1b && !(i1b && j1b && v1b)) || (nnz2 && !(i2 && j2 && k2 && l2 && v2)))
    return fail(c, ORDM_ERR_INVALID, "null sparse RDM arrays");
  CU(c, cudaSetDevice(c->device));
  std::vector<double> A1a, A1b;
  if (!ordm::sparse_opdm_to_dense(nact, nnz1a, i1a, j1a, v1a, upper_packed != 0, A1a) ||
      !ordm::sparse_opdm_to_dense(nact, nnz1b, i1b, j1b, v1b, upper_packed != 0, A1b))
    return fail(c, ORDM_ERR_INVALID, "sparse 1-RDM: index out of range [0,%d)%s", nact, upper_packed ? " or below the diagonal of an upper-packed list" : "");
  std::vector<double> Dt;
  if (!ordm::fold_tpdm_sparse(nact, nnz2, i2, j2, k2, l2, v2, upper_packed != 0, Dt))
    return fail(c, ORDM_ERR_INVALID, "sparse 2-RDM: index out of range [0,%d)%s", nact, upper_packed ? " or not i<=j, k<=l in an upper-packed list" : "");
  c->core_form = ncore > 0;
  int rc = install_rdm1(c, ncore, nact, A1a.data(), A1b.data());
  if (rc) return rc;
  c->h_Dt.swap(Dt);
  if ((rc = configure_k2(c))) return rc;
  c->have_rdm2 = true;
  return ORDM_OK;
}

int ordm_rdm2ab_sparse_to_dense(int n, size_t nnz, const int* i2, const int* j2, const int* k2, const int* l2, const double* v2,
                                int upper_packed, double* D2) {
  if (n <= 0 || !D2 || (nnz && !(i2 && j2 && k2 && l2 && v2))) return fail(nullptr, ORDM_ERR_INVALID, "bad sparse 2-RDM");
  if (!ordm::sparse_tpdm_to_dense(n, nnz, i2, j2, k2, l2, v2, upper_packed != 0, D2))
    return fail(nullptr, ORDM_ERR_INVALID, "sparse 2-RDM: index out of range [0,%d)%s", n, upper_packed ? " or not i<=j, k<=l in an upper-packed list" : "");
  return ORDM_OK;
}

// ---- stages -------------------------------------------------------------------------------
int ordm_build_rho(ordm_ctx* c) {
  CHECK_CTX(c);
  CU(c, cudaSetDevice(c->device));
  int rc = check_inputs(c, false);
  if (rc) return rc;
  if ((rc = ensure_vectors(c))) return rc;
  int launches = 0;
  if ((rc = run_k1(c, c->g, 0, &launches))) return rc;
  CU(c, cudaMemcpyAsync(c->h_res, c->d_res, ordm::RES_COUNT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < 3; ++i) c->last_n[i] = c->h_res[ordm::RES_N + i];
  c->rho_done = true;
  c->pi_done = c->xc_done = false;
  return ORDM_OK;
}

int ordm_build_pi(ordm_ctx* c) {
  CHECK_CTX(c);
  CU(c, cudaSetDevice(c->device));
  int rc = check_inputs(c, true);
  if (rc) return rc;
  if ((c->opts & ORDM_OPT_PI_GRADIENT) && !c->g.gga)
    return fail(c, ORDM_ERR_STATE, "ORDM_OPT_PI_GRADIENT needs the orbital gradients (phi_x, phi_y, phi_z)");
  if (c->core_form && !c->rho_done) return fail(c, ORDM_ERR_STATE, "build_pi with core orbitals needs build_rho first (closed-form core terms)");
  if ((rc = ensure_vectors(c))) return rc;
  int launches = 0;
  if ((rc = run_k2(c, c->g, &launches))) return rc;
  if (c->k2_i8 && !want_pig(c, c->g.gga)) {   // the int8 kernel ran: consult its time-out flag and its error guard
    int* h = reinterpret_cast<int*>(c->h_res + 12);
    CU(c, cudaMemcpyAsync(h, c->d_i8_fail, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    if (h[0] || h[1]) CU(c, cudaMemsetAsync(c->d_i8_fail, 0, 2 * sizeof(int), c->stream));
    if (h[0]) return fail(c, ORDM_ERR_CUDA, "int8 on-top kernel: %d CTA(s) timed out waiting for a tensor-core completion barrier", h[0]);
    if (h[1]) {   // flagged points: Pi again from the FP64 tensor-pipe kernels
      c->guard_points = (uint32_t)h[1]; c->i8_blocked = true; c->k2_i8 = false;
      if ((rc = run_k2(c, c->g, &launches))) return rc;
    }
  }
  CU(c, cudaStreamSynchronize(c->stream));
  c->pi_done = true;
  c->pig_done = want_pig(c, c->g.gga);
  c->xc_done = false;
  return ORDM_OK;
}

static int staged_xc(ordm_ctx* c, bool ft) {
  if (!c->rho_done || !c->pi_done) return fail(c, ORDM_ERR_STATE, "build_R/translate need build_rho and build_pi first");
  if (ft && c->g.gga && !c->pig_done)
    return fail(c, ORDM_ERR_STATE, "fully_translate with orbital gradients needs grad Pi: set ORDM_OPT_PI_GRADIENT (or ORDM_OPT_FULLY_TRANSLATED) before build_pi");
  if (c->xc_done && c->xc_ft == ft) return ORDM_OK;
  int launches = 0;
  if (c->inj_grad && c->g.gga) {   // the caller replaced spin-gradient components: rebuild the total gradient K3 reads
    double* tot[3] = {c->g.gx, c->g.gy, c->g.gz};
    for (int k = 0; k < 3; ++k) {
      const double *a = c->g.vec[ORDM_VEC_RHOA_X + k], *b = c->g.vec[ORDM_VEC_RHOB_X + k];
      if (!a || !b) return fail(c, ORDM_ERR_STATE, "translate after set_rhoa_x/...: both spin components of every gradient direction must exist (ORDM_OPT_DIAGNOSTICS)");
      if (c->g.npts) k_sum2<<<(unsigned)((c->g.npts + 255) / 256), 256, 0, c->stream>>>(tot[k], a, b, c->g.npts);
    }
    CU(c, cudaGetLastError());
    c->inj_grad = false;
  }
  int rc = run_k3(c, c->g, c->g.gga ? ORDM_FUNC_PBE : ORDM_FUNC_SVWN, ft, 0, &launches, false, c->inj_R ? c->g.vec[ORDM_VEC_R] : nullptr);
  if (rc) return rc;
  CU(c, cudaMemcpyAsync(c->h_res, c->d_res, ordm::RES_COUNT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < 3; ++i) c->last_trn[i] = c->h_res[ordm::RES_TRN + i];
  c->xc_done = true;
  c->xc_ft = ft;
  return ORDM_OK;
}

int ordm_build_R(ordm_ctx* c) {  // R is the same vector under both translations
  CHECK_CTX(c);
  CU(c, cudaSetDevice(c->device));
  if (c->inj_R) { c->inj_R = false; c->xc_done = false; }   // MCPDFT::build_R overwrites whatever set_R installed
  if (c->xc_done) return ORDM_OK;
  return staged_xc(c, (c->opts & ORDM_OPT_FULLY_TRANSLATED) != 0 && (!c->g.gga || c->pig_done));
}
int ordm_translate(ordm_ctx* c) {
  CHECK_CTX(c);
  CU(c, cudaSetDevice(c->device));
  return staged_xc(c, false);
}
int ordm_fully_translate(ordm_ctx* c) {
  CHECK_CTX(c);
  CU(c, cudaSetDevice(c->device));
  return staged_xc(c, true);
}

int ordm_integrated_densities(ordm_ctx* c, double n[3], double trn[3]) {
  CHECK_CTX(c);
  for (int i = 0; i < 3; ++i) {
    if (n) n[i] = c->last_n[i];
    if (trn) trn[i] = c->last_trn[i];
  }
  return ORDM_OK;
}

int ordm_functional(ordm_ctx* c, int term, size_t npts, const double* ra, const double* rb, const double* saa,
                    const double* sab, const double* sbb, double* e_out) {
  CHECK_CTX(c);
  if (term < 0 || term > 9 || !e_out || (npts && (!ra || !rb))) return fail(c, ORDM_ERR_INVALID, "bad functional call");
  if (!c->g.w || c->g.npts != npts) return fail(c, ORDM_ERR_STATE, "functional: set the %zu-point grid first (ordm_set_grid)", npts);
  const bool needs_ss = term == ORDM_EX_PBE || term == ORDM_EX_REVPBE || term == ORDM_EX_B88 || term == ORDM_EC_B88_OP || term == ORDM_EX_WPBE;
  const bool needs_all = term == ORDM_EC_PBE || term == ORDM_EC_LYP;
  if ((needs_ss && (!saa || !sbb)) || (needs_all && (!saa || !sab || !sbb)))
    return fail(c, ORDM_ERR_INVALID, "functional: missing sigma vectors");
  CU(c, cudaSetDevice(c->device));
  const size_t bytes = std::max<size_t>(npts, 1) * sizeof(double);
  double* d[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  const double* h[5] = {ra, rb, saa, sab, sbb};
  int rc = ORDM_OK;
  for (int i = 0; i < 5 && rc == ORDM_OK; ++i) {
    if (!h[i]) continue;
    if (cudaMalloc(&d[i], bytes) != cudaSuccess) { rc = fail(c, ORDM_ERR_CUDA, "out of device memory"); break; }
    if (npts && cudaMemcpyAsync(d[i], h[i], npts * sizeof(double), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) rc = fail(c, ORDM_ERR_CUDA, "H2D copy failed");
  }
  if (rc == ORDM_OK) {
    k3::FuncParams p{d[0], d[1], d[2], d[3], d[4], c->g.w, npts, nullptr, c->omega};
    const int grid = k3_grid(c, npts);
    rc = ensure_partials(c, (size_t)c->sm_count * 16 * 8);
    if (rc == ORDM_OK) {
      p.partial = c->d_partial;
      switch (term) {
        case 0: k3::k3_functional<0><<<grid, k3::THREADS, 0, c->stream>>>(p); break;
        case 1: k3::k3_functional<1><<<grid, k3::THREADS, 0, c->stream>>>(p); break;
        case 2: k3::k3_functional<2><<<grid, k3::THREADS, 0, c->stream>>>(p); break;
        case 3: k3::k3_functional<3><<<grid, k3::THREADS, 0, c->stream>>>(p); break;
        case 4: k3::k3_functional<4><<<grid, k3::THREADS, 0, c->stream>>>(p); break;
        case 5: k3::k3_functional<5><<<grid, k3::THREADS, 0, c->stream>>>(p); break;
        case 6: k3::k3_functional<6><<<grid, k3::THREADS, 0, c->stream>>>(p); break;
        case 7: k3::k3_functional<7><<<grid, k3::THREADS, 0, c->stream>>>(p); break;
        case 8: k3::k3_functional<8><<<grid, k3::THREADS, 0, c->stream>>>(p); break;
        default: k3::k3_functional<9><<<grid, k3::THREADS, 0, c->stream>>>(p); break;
      }
      ordm::k4_finalize<1><<<1, 256, 0, c->stream>>>(c->d_partial, grid, c->d_res, 8, 0);
      cudaError_t e = cudaMemcpyAsync(c->h_res, c->d_res + 8, sizeof(double), cudaMemcpyDeviceToHost, c->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
      if (e != cudaSuccess) rc = fail(c, ORDM_ERR_CUDA, "functional kernel failed: %s", cudaGetErrorString(e));
      else *e_out = c->h_res[0];
    }
  }
  cudaStreamSynchronize(c->stream);
  for (auto& p : d) dfree(p);
  return rc;
}

// MCPDFTSolver::polyradical_analysis (src/libpsi4/mcpdft_solver.cc:3001-3240) on the resident grid and 1-RDMs
int ordm_polyradical(ordm_ctx* c, double traces[3], double* U_r, double* D_r, double* N_r, size_t npts) {
  CHECK_CTX(c);
  if (!traces) return fail(c, ORDM_ERR_INVALID, "null traces");
  CU(c, cudaSetDevice(c->device));
  int rc = check_inputs(c, false);
  if (rc) return rc;
  if ((U_r || D_r || N_r) && npts != c->g.npts) return fail(c, ORDM_ERR_INVALID, "vector length %zu != grid size %zu", npts, c->g.npts);
  if (c->nact > k1::MAX_ACT) return fail(c, ORDM_ERR_UNSUPPORTED, "polyradical analysis: %d active orbitals (limit %d)", c->nact, k1::MAX_ACT);
  const int na = c->nact;
  std::vector<double> udn((size_t)3 * na * na);
  for (int t = 0; t < na; ++t)
    for (int u = 0; u < na; ++u) {   // element (t, u) of the column-major inputs, row-major here
      const double pop = c->h_A1a[t + (size_t)u * na] + c->h_A1b[t + (size_t)u * na];   // mcpdft_solver.cc:3082-3084, 3102
      udn[(size_t)t * na + u] = 1.0 - std::fabs(1.0 - pop);                              // :3104
      udn[(size_t)na * na + (size_t)t * na + u] = pop * (2.0 - pop);                     // :3105
      udn[(size_t)2 * na * na + (size_t)t * na + u] = pop * pop * (2.0 - pop) * (2.0 - pop);   // :3106
    }
  double* d_udn = nullptr;
  double* d_out[3] = {nullptr, nullptr, nullptr};
  double* h_out[3] = {U_r, D_r, N_r};
  auto cleanup = [&]() { cudaStreamSynchronize(c->stream); dfree(d_udn); for (auto& q : d_out) dfree(q); };
#define CUP(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail(c, ORDM_ERR_CUDA, "CUDA error '%s' in %s", cudaGetErrorString(e_), #call); } } while (0)
  CUP(cudaMalloc(&d_udn, udn.size() * sizeof(double)));
  CUP(cudaMemcpyAsync(d_udn, udn.data(), udn.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  for (int k = 0; k < 3; ++k)
    if (h_out[k]) CUP(cudaMalloc(&d_out[k], std::max<size_t>(c->g.ld, 1) * sizeof(double)));
  k1::PolyParams p{};
  p.phi_act = c->g.phi + (size_t)c->ncore * c->g.ld; p.w = c->g.w; p.udn = d_udn;
  p.ld = c->g.ld; p.npts = c->g.npts; p.nact = na;
  p.U = d_out[0]; p.D = d_out[1]; p.N = d_out[2];
  const int grid = k1_grid(c, c->g.npts);
  if ((rc = ensure_partials(c, (size_t)c->sm_count * 16 * 8))) { cleanup(); return rc; }
  p.partial = c->d_partial;
  switch ((int)round_up((size_t)std::max(na, 1), 4)) {
#define PCASE(N) case N: k1::k1_polyradical<N><<<grid, k1::THREADS, 0, c->stream>>>(p); break;
    PCASE(4) PCASE(8) PCASE(12) PCASE(16) PCASE(20) PCASE(24) PCASE(28) default: k1::k1_polyradical<32><<<grid, k1::THREADS, 0, c->stream>>>(p); break;
#undef PCASE
  }
  CUP(cudaGetLastError());
  ordm::k4_finalize<3><<<1, 256, 0, c->stream>>>(c->d_partial, grid, c->d_res, 8, 0);
  CUP(cudaGetLastError());
  CUP(cudaMemcpyAsync(c->h_res, c->d_res + 8, 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  for (int k = 0; k < 3; ++k)
    if (h_out[k] && npts) CUP(cudaMemcpyAsync(h_out[k], d_out[k], npts * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUP(cudaStreamSynchronize(c->stream));
#undef CUP
  for (int k = 0; k < 3; ++k) traces[k] = c->h_res[k];
  cleanup();
  return ORDM_OK;
}

// Energy assembly of MCPDFTSolver::compute_energy (src/libpsi4/mcpdft_solver.cc:849-925); host arithmetic only
int ordm_hybrid_energy(int method, double lambda, const ordm_hybrid_inputs* in, double ex, double ec, ordm_hybrid_result* out) {
  if (!in || !out || method < ORDM_METHOD_MCPDFT || method > ORDM_METHOD_RS1DH_MCPDFT) return fail(nullptr, ORDM_ERR_INVALID, "bad hybrid-energy request");
  const double lmbd = method == ORDM_METHOD_MCPDFT ? 0.0 : lambda;                     // :717-737: lambda is read for every method but MCPDFT
  out->lambda = lmbd;
  out->one_electron = in->e_nuclear_attraction + in->e_kinetic;                      // :830
  out->two_electron = in->e_reference - in->e_nuclear_repulsion - out->one_electron; // :852
  out->wf_contribution = out->one_electron + lmbd * out->two_electron;               // :918
  out->dft_contribution = (1.0 - lmbd) * (in->e_coulomb + ex) + (1.0 - lmbd * lmbd) * ec;   // :919
  out->e_total = out->wf_contribution + out->dft_contribution + in->e_nuclear_repulsion;    // :920
  if (method == ORDM_METHOD_1DH_MCPDFT) out->e_total += (lmbd * in->e_hf_exchange) + (lmbd * lmbd * in->e_mp2_correlation);   // :921-922
  return ORDM_OK;
}

// ---- the path in one call -----------------------------------------------------------------
static int energy_enqueue(ordm_ctx* c, int functional, double eclass);
static int energy_finish(ordm_ctx* c, ordm_result* out);
static int energy_resident(ordm_ctx* c, int functional, double eclass, ordm_result* out) {
  int rc = energy_enqueue(c, functional, eclass);
  return rc ? rc : energy_finish(c, out);
}
int ordm_energy(ordm_ctx* c, int functional, double eclass, ordm_result* out) {
  CHECK_CTX(c);
  if (!out) return fail(c, ORDM_ERR_INVALID, "null result");
  if (c->pend.active) return fail(c, ORDM_ERR_STATE, "a step enqueued by ordm_energy_async is still pending: collect it with ordm_energy_wait first");
  int rc = energy_resident(c, functional, eclass, out);
  if (rc == RETRY_FP64) {   // the int8 kernel's guard flagged points: same call again, Pi from the FP64 tensor-pipe kernels
    rc = energy_resident(c, functional, eclass, out);
    if (rc == ORDM_OK) out->pi_guard_points = c->guard_points;
  }
  return rc;
}
// Stream-ordered form: ordm_energy_async enqueues the whole step (kernels, all-reduce, D2H read of the scalars into the
// context's pinned result block) and returns; ordm_energy_wait blocks for the LAST enqueued step and returns its
// result.  Several steps may be enqueued back to back (an SCF driver that pipelines, a benchmark loop): the host then
// runs ahead of the GPU and the per-step launch latency disappears from the device timeline.
int ordm_energy_async(ordm_ctx* c, int functional, double eclass) {
  CHECK_CTX(c);
  return energy_enqueue(c, functional, eclass);
}
int ordm_energy_wait(ordm_ctx* c, ordm_result* out) {
  CHECK_CTX(c);
  if (!out) return fail(c, ORDM_ERR_INVALID, "null result");
  if (!c->pend.active) return fail(c, ORDM_ERR_STATE, "ordm_energy_wait without a pending ordm_energy_async");
  const int functional = c->pend.functional;
  const double eclass = c->pend.eclass;
  int rc = energy_finish(c, out);
  if (rc == RETRY_FP64) {   // some enqueued step tripped the int8 guard: the last one is repeated with the FP64 kernels
    rc = energy_resident(c, functional, eclass, out);
    if (rc == ORDM_OK) out->pi_guard_points = c->guard_points;
  }
  return rc;
}
static int energy_enqueue(ordm_ctx* c, int functional, double eclass) {
  CU(c, cudaSetDevice(c->device));
  int rc = check_inputs(c, true);
  if (rc) return rc;
  if ((c->opts & ORDM_OPT_PI_GRADIENT) && !c->g.gga)
    return fail(c, ORDM_ERR_STATE, "ORDM_OPT_PI_GRADIENT needs the orbital gradients (phi_x, phi_y, phi_z)");
  if ((rc = ensure_vectors(c))) return rc;
  const bool ft = (c->opts & ORDM_OPT_FULLY_TRANSLATED) != 0;
  c->inj_R = c->inj_grad = false;   // mcpdft_energy rebuilds every vector from the inputs (energy.cc:49-58)
  int launches = 0;
  int grid1 = -1, grid3 = 0;   // rows of K1's / K3's block sums left for the final kernel (grid1 < 0: K1 finalised itself)
  // Overlapped mode (default): the tensor-bound K2 is launched first with a trimmed register
  // footprint, and the HBM-bound K1 streams underneath it on a second stream (one K1 CTA stays
  // co-resident per SM); K3 joins both and completes Pi = Pi_act + Pi_core.
  // (only when K2 leaves a comfortable part of the SM's shared memory free: with 26 active orbitals K2
  //  takes 219 KB of the 228 and, although a K1 CTA nominally still fits, it was measured not to
  //  co-run -- K1 then queues behind the slightly slower lean K2: 93.6 vs 92.7 ms on config 5)
  // With many core orbitals only their FMA-light, byte-heavy sums run underneath K2 (k1_core_light);
  // the active-space part of K1 follows on the main stream once both are done.
  const bool split_ok = c->core_form && c->ncore >= c->nact && c->g.core[0] && !c->no_split;
  // The int8 K2 owns the SM's shared memory and TMEM but leaves 17 K registers: only the split form fits beside it
  // (one deep-unrolled core CTA per SM); without core orbitals to stream the stages run back to back.
  // A K2 that takes (almost) the whole shared memory -- 26 active orbitals: 219 KB -- still leaves room for the deep
  // core-orbital kernel, which uses none: it is placed first (as beside the int8 K2) and streams the core columns
  // while K2 owns the tensor pipe.
  const bool k2_fills_smem = c->k2_smem + 32768 > (size_t)c->max_smem_optin;
  const bool deep_beside_dmma = !c->k2_i8 && k2_fills_smem && split_ok && c->big_overlap;
  const bool overlap = !c->serial_stages && !(c->opts & ORDM_OPT_SERIAL_STAGES) && c->k2_ws && !want_pig(c, c->g.gga) &&
                       (!k2_fills_smem || deep_beside_dmma) && (!c->k2_i8 || (split_ok && !c->i8_no_overlap));
  const bool split = overlap && split_ok;
  const bool core_first = c->k2_i8 || deep_beside_dmma;   // the deep core CTA takes its slot on every SM before K2 arrives
  CU(c, cudaEventRecord(c->ev[0], c->stream));
  if (overlap) {
    CU(c, cudaEventRecord(c->ev_fork, c->stream));
    CU(c, cudaStreamWaitEvent(c->side_stream, c->ev_fork, 0));
    // Launch order: the DMMA K2 goes first and the light core kernel fills the registers it leaves; the int8 K2
    // takes one CTA slot per SM whatever is there, so the single deep core CTA per SM is placed first.
    if (!core_first) {
      if ((rc = run_k2(c, c->g, &launches, /*lean=*/true))) return rc;
      CU(c, cudaEventRecord(c->ev[2], c->stream));
    }
    CU(c, cudaEventRecord(c->ev_k1[0], c->side_stream));
    const bool act_under = split && c->k2_i8 && c->k1act_under;   // the active-space part too beside the int8 K2
    if (split) {
      if ((rc = run_k1_core_light(c, c->g, &launches, c->side_stream, /*deep=*/core_first))) return rc;
      if (act_under && (rc = run_k1(c, c->g, 0, &launches, c->side_stream, /*core_ready=*/true, /*under_k2=*/true))) return rc;
    } else if ((rc = run_k1(c, c->g, 0, &launches, c->side_stream))) return rc;
    CU(c, cudaEventRecord(c->ev_k1[1], c->side_stream));
    if (core_first) {
      if ((rc = run_k2(c, c->g, &launches, /*lean=*/true))) return rc;
      CU(c, cudaEventRecord(c->ev[2], c->stream));
    }
    CU(c, cudaEventRecord(c->ev_join, c->side_stream));
    CU(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    CU(c, cudaEventRecord(c->ev[1], c->stream));  // = K2 and the side-stream part of K1 done
    if (split && !act_under && (rc = run_k1(c, c->g, 0, &launches, nullptr, /*core_ready=*/true, false, &grid1))) return rc;
    CU(c, cudaEventRecord(c->ev[6], c->stream));
    if ((rc = run_k3(c, c->g, functional, ft, 0, &launches, /*add_core=*/true, nullptr, &grid3))) return rc;
  } else {
    if ((rc = run_k1(c, c->g, 0, &launches, nullptr, false, false, &grid1))) return rc;
    CU(c, cudaEventRecord(c->ev[1], c->stream));
    if ((rc = run_k2(c, c->g, &launches))) return rc;
    CU(c, cudaEventRecord(c->ev[2], c->stream));
    if ((rc = run_k3(c, c->g, functional, ft, 0, &launches, false, nullptr, &grid3))) return rc;
  }
  CU(c, cudaEventRecord(c->ev[3], c->stream));
  // one launch ends the step: final passes over K1's and K3's block sums + (peer-mapped communicator) the all-reduce
  bool host_direct = false;
  {
    const bool peer = c->comm && c->peer_ok;
    ordm::PeerXchg px = peer ? c->px : ordm::PeerXchg{};
    if (!peer) px.nranks = 1;
    // (without the NCCL fall-back the kernel also delivers the result: scalars + status words into the pinned block)
    host_direct = (!c->comm || peer) && !c->res_by_copy;
    double* host_out = host_direct ? c->h_res : nullptr;
    const int* status = (c->k2_i8 && c->d_i8_fail) ? c->d_i8_fail : nullptr;
    if (grid1 < 0) {   // K1's final pass already ran on its own stream (overlap without split, or act_under): fold nothing here
      static const double* none = nullptr;
      ordm::k4_final_all<<<1, 256, 0, c->stream>>>(none, -1, c->d_partial3, grid3, c->d_res, px, peer ? ++c->xepoch : 0ull, c->d_xerr, host_out, status);
    } else {
      ordm::k4_final_all<<<1, 256, 0, c->stream>>>(c->d_partial, grid1, c->d_partial3, grid3, c->d_res, px, peer ? ++c->xepoch : 0ull, c->d_xerr, host_out, status);
    }
    CU(c, cudaGetLastError());
    launches += 1;
    if (c->comm && !peer) {   // NCCL fall-back
      int nr = g_nccl.AllReduce(c->d_res, c->d_res, ordm::RES_COUNT, /*ncclDouble*/ 8, /*ncclSum*/ 0, c->comm, c->stream);
      if (nr != 0) return fail(c, ORDM_ERR_NCCL, "ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(nr) : "?");
      launches += 1;
    }
  }
  CU(c, cudaEventRecord(c->ev[4], c->stream));
  if ((rc = fetch_enqueue(c, host_direct))) return rc;
  c->pend.active = true; c->pend.overlap = overlap; c->pend.split = split; c->pend.ft = ft;
  c->pend.functional = functional; c->pend.launches = launches; c->pend.eclass = eclass;
  return ORDM_OK;
}

static int energy_finish(ordm_ctx* c, ordm_result* out) {
  const bool overlap = c->pend.overlap, split = c->pend.split, ft = c->pend.ft;
  const int functional = c->pend.functional;
  c->pend.active = false;
  int rc;
  if ((rc = fetch_finish(c, c->pend.eclass, out, c->g.npts, c->pend.launches))) return rc;
  float ms;
  if (overlap) {  // K1 (or its core part) and K2 ran concurrently: each is timed on its own stream
    CU(c, cudaEventElapsedTime(&ms, c->ev_k1[0], c->ev_k1[1])); out->ms_density = ms;
    CU(c, cudaEventElapsedTime(&ms, c->ev[1], c->ev[6])); out->ms_density += ms;  // + the active-space part (split mode)
    CU(c, cudaEventElapsedTime(&ms, c->ev[0], c->ev[2])); out->ms_pi = ms;
    CU(c, cudaEventElapsedTime(&ms, c->ev[6], c->ev[3])); out->ms_xc = ms;
    out->overlapped = split ? 2 : 1;
  } else {
    CU(c, cudaEventElapsedTime(&ms, c->ev[0], c->ev[1])); out->ms_density = ms;
    CU(c, cudaEventElapsedTime(&ms, c->ev[1], c->ev[2])); out->ms_pi = ms;
    CU(c, cudaEventElapsedTime(&ms, c->ev[2], c->ev[3])); out->ms_xc = ms;
  }
  CU(c, cudaEventElapsedTime(&ms, c->ev[3], c->ev[4])); out->ms_reduce = ms;
  CU(c, cudaEventElapsedTime(&ms, c->ev[0], c->ev[5])); out->ms_total = ms;
  c->rho_done = c->pi_done = true;
  c->pig_done = want_pig(c, c->g.gga);
  c->xc_done = false;  // R/translated vectors belong to `functional`; staged getters recompute
  if ((functional != ORDM_FUNC_SVWN) == c->g.gga) { c->xc_done = true; c->xc_ft = ft; }  // R / translated vectors do not depend on which GGA
  return ORDM_OK;
}

static int energy_streamed(ordm_ctx* c, int functional, double eclass, size_t npts, int n, size_t ld, const double* w,
                           const double* phi, const double* px, const double* py, const double* pz, size_t batch, ordm_result* out);
int ordm_energy_host(ordm_ctx* c, int functional, double eclass, size_t npts, int n, size_t ld, const double* w,
                     const double* phi, const double* px, const double* py, const double* pz, size_t batch,
                     ordm_result* out) {
  CHECK_CTX(c);
  if (!out || !w || !phi || n <= 0 || ld < npts) return fail(c, ORDM_ERR_INVALID, "bad host buffers");
  // the host buffers are new inputs as far as the int8 guard is concerned
  c->i8_blocked = false; c->k2_i8 = c->k2_i8_cfg;
  int rc = energy_streamed(c, functional, eclass, npts, n, ld, w, phi, px, py, pz, batch, out);
  if (rc == RETRY_FP64) {
    rc = energy_streamed(c, functional, eclass, npts, n, ld, w, phi, px, py, pz, batch, out);
    if (rc == ORDM_OK) out->pi_guard_points = c->guard_points;
  }
  return rc;
}
static int energy_streamed(ordm_ctx* c, int functional, double eclass, size_t npts, int n, size_t ld, const double* w,
                           const double* phi, const double* px, const double* py, const double* pz, size_t batch, ordm_result* out) {
  const bool gga = px || py || pz;
  if (gga && !(px && py && pz)) return fail(c, ORDM_ERR_INVALID, "give all of phi_x, phi_y, phi_z or none");
  if (!c->have_rdm1 || !c->have_rdm2) return fail(c, ORDM_ERR_STATE, "RDMs not set");
  if (c->ncore + c->nact != n) return fail(c, ORDM_ERR_STATE, "RDM dimension %d+%d does not match n=%d", c->ncore, c->nact, n);
  CU(c, cudaSetDevice(c->device));
  if (batch == 0) batch = 1u << 18;
  batch = round_up(std::min(batch, std::max<size_t>(npts, 1)), LD_ALIGN);
  // (re)allocate the two staging sets
  if (batch > c->stage_pts || n > c->stage_n || (gga && !c->stage_gga)) {
    for (auto& s : c->stage) {
      dfree(s.w);
      for (auto& p : s.phi) dfree(p);
      for (auto& p : s.vec) dfree(p);
      CU(c, cudaMalloc(&s.w, batch * sizeof(double)));
      for (int k = 0; k < (gga ? 4 : 1); ++k) {
        CU(c, cudaMalloc(&s.phi[k], batch * (size_t)n * sizeof(double)));
        CU(c, cudaMemset(s.phi[k], 0, batch * (size_t)n * sizeof(double)));
      }
      for (auto& p : s.vec) CU(c, cudaMalloc(&p, batch * sizeof(double)));
      if (!s.copied) CU(c, cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
      if (!s.done) CU(c, cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    }
    c->stage_pts = batch; c->stage_n = n; c->stage_gga = gga;
  }
  int launches = 0;
  CU(c, cudaMemsetAsync(c->d_res, 0, 16 * sizeof(double), c->stream));
  CU(c, cudaEventRecord(c->ev[0], c->stream));
  const double* src[4] = {phi, px, py, pz};
  size_t ib = 0;
  for (size_t p0 = 0; p0 < npts; p0 += batch, ++ib) {
    ordm_ctx::Stage& s = c->stage[ib & 1];
    const size_t m = std::min(batch, npts - p0);
    if (ib >= 2) CU(c, cudaStreamWaitEvent(c->copy_stream, s.done, 0));  // staging set is free again
    CU(c, cudaMemcpyAsync(s.w, w + p0, m * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
    for (int k = 0; k < (gga ? 4 : 1); ++k)
      CU(c, cudaMemcpy2DAsync(s.phi[k], c->stage_pts * sizeof(double), src[k] + p0, ld * sizeof(double), m * sizeof(double), (size_t)n, cudaMemcpyHostToDevice, c->copy_stream));
    CU(c, cudaEventRecord(s.copied, c->copy_stream));
    CU(c, cudaStreamWaitEvent(c->stream, s.copied, 0));
    GridView v;
    v.npts = m; v.ld = c->stage_pts; v.n = n; v.gga = gga;
    v.w = s.w; v.phi = s.phi[0]; v.phix = s.phi[1]; v.phiy = s.phi[2]; v.phiz = s.phi[3];
    v.vec[ORDM_VEC_RHO] = s.vec[0]; v.vec[ORDM_VEC_PI] = s.vec[1];
    v.gx = s.vec[2]; v.gy = s.vec[3]; v.gz = s.vec[4]; v.pi_core = s.vec[5];
    for (int q = 0; q < 4; ++q) v.core[q] = s.vec[6 + q];
    for (int q = 0; q < 3; ++q) { v.vec[ORDM_VEC_PI_X + q] = s.vec[10 + q]; v.pcg[q] = s.vec[13 + q]; }
    int rc;
    if ((rc = run_k1(c, v, 1, &launches))) return rc;
    if ((rc = run_k2(c, v, &launches))) return rc;
    if ((rc = run_k3(c, v, functional, (c->opts & ORDM_OPT_FULLY_TRANSLATED) != 0, 1, &launches))) return rc;
    CU(c, cudaEventRecord(s.done, c->stream));
  }
  CU(c, cudaEventRecord(c->ev[3], c->stream));
  int rc = all_reduce_results(c, &launches);
  if (rc) return rc;
  CU(c, cudaEventRecord(c->ev[4], c->stream));
  rc = fetch_result(c, eclass, out, npts, launches);
  if (rc) return rc;
  float ms;
  CU(c, cudaEventElapsedTime(&ms, c->ev[0], c->ev[5])); out->ms_total = ms;
  CU(c, cudaEventElapsedTime(&ms, c->ev[3], c->ev[4])); out->ms_reduce = ms;
  return ORDM_OK;
}

// MCPDFT::set_rhoa ... set_tr_sigma_bb, set_pi, set_pi_x/y/z, set_R (mcpdft.h:125-153): the caller's vector replaces
// the engine's and is what the following stages and getters see.
int ordm_set_vector(ordm_ctx* c, int which, const double* host_in, size_t npts) {
  CHECK_CTX(c);
  if (which < 0 || which >= ORDM_VEC_COUNT || (npts && !host_in)) return fail(c, ORDM_ERR_INVALID, "bad vector");
  if (which == ORDM_VEC_W) return ordm_set_grid(c, npts, host_in);
  if (!c->g.w || npts != c->g.npts) return fail(c, ORDM_ERR_INVALID, "vector length %zu != grid size %zu (ordm_set_grid first)", npts, c->g.npts);
  CU(c, cudaSetDevice(c->device));
  if (c->g.ld < round_up(std::max<size_t>(npts, 1), LD_ALIGN)) c->g.ld = round_up(std::max<size_t>(npts, 1), LD_ALIGN);  // no orbitals yet
  int rc = ensure_vectors(c);
  if (rc) return rc;
  double*& dst = c->g.vec[which];
  if (!dst) {   // energy-only mode keeps few vectors: an installed one is materialised on its own
    CU(c, cudaMalloc(&dst, c->g.ld * sizeof(double)));
    CU(c, cudaMemsetAsync(dst, 0, c->g.ld * sizeof(double), c->stream));
  }
  if (npts) CU(c, cudaMemcpyAsync(dst, host_in, npts * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  // the translated vectors are outputs of translate(): installing one makes the set caller-managed (the getters hand
  // back what is stored, as the reference's do) until the next translate; anything upstream invalidates them
  c->xc_done = which >= ORDM_VEC_TR_RHOA && which <= ORDM_VEC_TR_SIGMA_BB;
  if (which == ORDM_VEC_RHO) c->rho_done = true;         // what build_R / translate read (mcpdft.cc:276,297)
  else if (which == ORDM_VEC_PI) c->pi_done = true;      // the complete Pi (core part included)
  else if (which == ORDM_VEC_R) c->inj_R = true;
  else if (which >= ORDM_VEC_RHOA_X && which <= ORDM_VEC_RHOB_Z) c->inj_grad = true;
  else if (which >= ORDM_VEC_PI_X && which <= ORDM_VEC_PI_Z) c->pig_done = true;
  return ORDM_OK;
}

// ---- input files: the PySCF exporter's data.h5 schema (raw container, or HDF5 when built with it; rawio.h) ----------
struct ordm_file { ordm::RawFile f; };

int ordm_file_open(const char* path, ordm_file** out) {
  if (!path || !out) return fail(nullptr, ORDM_ERR_INVALID, "null path / output");
  *out = nullptr;
  ordm_file* h = new ordm_file();
  if (!h->f.open(path)) { const int rc = fail(nullptr, ORDM_ERR_INVALID, "%s", h->f.error.c_str()); delete h; return rc; }
  *out = h;
  return ORDM_OK;
}
void ordm_file_close(ordm_file* f) { delete f; }
int ordm_file_dataset(ordm_file* f, const char* name, int* dtype, int* ndim, uint64_t dims[4], const void** data) {
  if (!f || !name) return fail(nullptr, ORDM_ERR_INVALID, "null file / name");
  const ordm::RawDataset* d = f->f.find(name);
  if (!d) return fail(nullptr, ORDM_ERR_INVALID, "dataset %s not in the file", name);
  if (dtype) *dtype = d->dtype;
  if (ndim) *ndim = d->ndim;
  if (dims) for (int i = 0; i < 4; ++i) dims[i] = i < d->ndim ? d->dims[i] : 1;
  if (data) *data = d->data;
  return ORDM_OK;
}

int ordm_load_data(ordm_ctx* c, const char* path, unsigned flags, ordm_data_info* info) {
  CHECK_CTX(c);
  if (!path) return fail(c, ORDM_ERR_INVALID, "null path");
  CU(c, cudaSetDevice(c->device));
  ordm::RawFile f;
  if (!f.open(path)) return fail(c, ORDM_ERR_INVALID, "%s", f.error.c_str());
  ordm_data_info di;
  std::memset(&di, 0, sizeof(di));
  double v = 0.0;
  if (!f.scalar("/N/N_COR", v)) return fail(c, ORDM_ERR_INVALID, "%s: /N/N_COR missing", path);
  di.ncore = (int)v;
  if (!f.scalar("/N/N_CAS_ORB", v)) return fail(c, ORDM_ERR_INVALID, "%s: /N/N_CAS_ORB missing", path);
  di.nact = (int)v;
  if (f.scalar("/N/N_AO", v)) di.nao = (int)v;
  if (f.scalar("/N/N_MO", v)) di.nmo = (int)v;
  if (f.scalar("/E/E_CORE", v)) di.e_core = v;
  if (f.scalar("/E/E_HARTREE", v)) di.e_hartree = v;
  const int nocc = di.ncore + di.nact;
  if (di.ncore < 0 || di.nact <= 0) return fail(c, ORDM_ERR_INVALID, "%s: bad orbital space (%d core, %d active)", path, di.ncore, di.nact);
  // ---- grid: /GRIDS/W (libpyscf.py:654)
  const ordm::RawDataset* w = f.find("/GRIDS/W");
  if (!w || w->dtype != ordm::RAW_F64 || w->ndim != 1) return fail(c, ORDM_ERR_INVALID, "%s: /GRIDS/W missing or not a float64 vector", path);
  const size_t npts = (size_t)w->dims[0];
  di.npts = npts;
  int rc = ordm_set_grid(c, npts, (const double*)w->data);
  if (rc) return rc;
  // ---- orbitals: /PHI/PHI_MO[npts][nmo] (+ _X, _Y, _Z), or /PHI/PHI_AO[npts][nao] with /C[nao][nmo] (libpyscf.py:608,659-667)
  const char* suffix[4] = {"", "_X", "_Y", "_Z"};
  const ordm::RawDataset* mo[4]; const ordm::RawDataset* ao[4];
  for (int k = 0; k < 4; ++k) { mo[k] = f.find(std::string("/PHI/PHI_MO") + suffix[k]); ao[k] = f.find(std::string("/PHI/PHI_AO") + suffix[k]); }
  const ordm::RawDataset* Cm = f.find("/C");
  auto shaped = [&](const ordm::RawDataset* d) { return d && d->dtype == ordm::RAW_F64 && d->ndim == 2 && d->dims[0] == npts; };
  const bool have_mo = shaped(mo[0]) && (int)mo[0]->dims[1] >= nocc;
  const bool have_ao = shaped(ao[0]) && Cm && Cm->dtype == ordm::RAW_F64 && Cm->ndim == 2 && Cm->dims[0] == ao[0]->dims[1] && (int)Cm->dims[1] >= nocc;
  const bool use_ao = have_ao && ((flags & ORDM_LOAD_PREFER_AO) || !have_mo);
  if (!use_ao && !have_mo) return fail(c, ORDM_ERR_INVALID, "%s: neither /PHI/PHI_MO nor /PHI/PHI_AO + /C with %zu points and >= %d orbitals", path, npts, nocc);
  const ordm::RawDataset* const* phi = use_ao ? ao : mo;
  const bool gga = !(flags & ORDM_LOAD_NO_GRADIENTS) && shaped(phi[1]) && shaped(phi[2]) && shaped(phi[3]) &&
                   phi[1]->dims[1] == phi[0]->dims[1] && phi[2]->dims[1] == phi[0]->dims[1] && phi[3]->dims[1] == phi[0]->dims[1];
  di.gga = gga; di.used_ao = use_ao;
  if (use_ao) {
    const int nao = (int)ao[0]->dims[1], nmo_all = (int)Cm->dims[1];
    di.nao = nao; di.nmo = nmo_all;
    std::vector<double> Ccm((size_t)nao * nocc);   // /C is row-major [nao][nmo]: the occupied columns, column-major
    for (int i = 0; i < nocc; ++i)
      for (int mu = 0; mu < nao; ++mu) Ccm[(size_t)i * nao + mu] = ((const double*)Cm->data)[(size_t)mu * nmo_all + i];
    const double* src[4] = {(const double*)ao[0]->data, gga ? (const double*)ao[1]->data : nullptr, gga ? (const double*)ao[2]->data : nullptr,
                            gga ? (const double*)ao[3]->data : nullptr};
    if ((rc = ao_to_mo(c, npts, nao, (size_t)nao, src, gga, nocc, Ccm.data(), /*row_major=*/true))) return rc;
  } else {
    const int nmo_all = (int)mo[0]->dims[1];
    di.nmo = nmo_all;
    if ((rc = alloc_grid(c, npts, nocc, gga))) return rc;
    const size_t batch = std::max<size_t>(1024, ((size_t)1 << 24) / (size_t)nmo_all);   // rows per staging buffer (<= 128 MB)
    double* d_stage = nullptr;
    CU(c, cudaMalloc(&d_stage, batch * (size_t)nmo_all * sizeof(double)));
    for (int k = 0; k < (gga ? 4 : 1); ++k)
      for (size_t p0 = 0; p0 < npts; p0 += batch) {
        const size_t m = std::min(batch, npts - p0);
        cudaError_t e = cudaMemcpyAsync(d_stage, (const double*)mo[k]->data + p0 * nmo_all, m * (size_t)nmo_all * sizeof(double), cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) {
          dim3 grid((unsigned)((m + 31) / 32), (unsigned)((nocc + 31) / 32)), block(32, 8);
          k_rows_to_columns<<<grid, block, 0, c->stream>>>(c->d_phi[k] + p0, c->g.ld, d_stage, m, nmo_all, nocc);
          e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);   // the staging buffer is reused (and the source may be pageable)
        if (e != cudaSuccess) { dfree(d_stage); return fail(c, ORDM_ERR_CUDA, "loading /PHI/PHI_MO%s: %s", suffix[k], cudaGetErrorString(e)); }
      }
    dfree(d_stage);
  }
  // ---- RDMs of the active space: the upper-packed COO groups (libpyscf.py:246-463) or the dense blocks (:636-648)
  std::vector<int> i1a, j1a, i1b, j1b, q1, q2, q3, q4;
  const ordm::RawDataset *v1a = f.find("/SP_SYMM_D1/ACT_D1a_MO/VAL"), *v1b = f.find("/SP_SYMM_D1/ACT_D1b_MO/VAL"), *v2 = f.find("/SP_SYMM_D2/ACT_D2ab_MO/VAL");
  const bool sparse_ok = !(flags & ORDM_LOAD_DENSE_RDM) && v1a && v1b && v2 && v1a->dtype == ordm::RAW_F64 && v1b->dtype == ordm::RAW_F64 && v2->dtype == ordm::RAW_F64 &&
                         f.ints("/SP_SYMM_D1/ACT_D1a_MO/ROW_IDX", i1a) && f.ints("/SP_SYMM_D1/ACT_D1a_MO/COL_IDX", j1a) &&
                         f.ints("/SP_SYMM_D1/ACT_D1b_MO/ROW_IDX", i1b) && f.ints("/SP_SYMM_D1/ACT_D1b_MO/COL_IDX", j1b) &&
                         f.ints("/SP_SYMM_D2/ACT_D2ab_MO/DIM1_IDX", q1) && f.ints("/SP_SYMM_D2/ACT_D2ab_MO/DIM2_IDX", q2) &&
                         f.ints("/SP_SYMM_D2/ACT_D2ab_MO/DIM3_IDX", q3) && f.ints("/SP_SYMM_D2/ACT_D2ab_MO/DIM4_IDX", q4) &&
                         i1a.size() == v1a->count() && j1a.size() == i1a.size() && i1b.size() == v1b->count() && j1b.size() == i1b.size() &&
                         q1.size() == v2->count() && q2.size() == q1.size() && q3.size() == q1.size() && q4.size() == q1.size();
  di.used_sparse_rdm = sparse_ok;
  if (sparse_ok) {
    rc = ordm_set_active_space_sparse(c, di.ncore, di.nact, i1a.size(), i1a.data(), j1a.data(), (const double*)v1a->data,
                                      i1b.size(), i1b.data(), j1b.data(), (const double*)v1b->data,
                                      q1.size(), q1.data(), q2.data(), q3.data(), q4.data(), (const double*)v2->data, /*upper_packed=*/1);
  } else {
    const ordm::RawDataset *d1a = f.find("/D/D1/ACT_D1A_MO"), *d1b = f.find("/D/D1/ACT_D1B_MO"), *d2 = f.find("/D/D2/ACT_D2AB_MO");
    const size_t na = (size_t)di.nact;
    if (!d1a || !d1b || !d2 || d1a->count() != na * na || d1b->count() != na * na || d2->count() != na * na * na * na ||
        d1a->dtype != ordm::RAW_F64 || d1b->dtype != ordm::RAW_F64 || d2->dtype != ordm::RAW_F64)
      return fail(c, ORDM_ERR_INVALID, "%s: no usable active-space RDMs (/SP_SYMM_D1,2/ACT_* or /D/D1/ACT_D1A_MO, ACT_D1B_MO, /D/D2/ACT_D2AB_MO)", path);
    // dm2ab[i][j][k][l] multiplies phi_i phi_j (alpha) phi_k phi_l (beta): every element once, as a Psi4-style list
    std::vector<int> ii(d2->count()), jj(d2->count()), kk(d2->count()), ll(d2->count());
    size_t e = 0;
    for (int i = 0; i < di.nact; ++i) for (int j = 0; j < di.nact; ++j) for (int k = 0; k < di.nact; ++k) for (int l = 0; l < di.nact; ++l, ++e) { ii[e] = i; jj[e] = j; kk[e] = k; ll[e] = l; }
    std::vector<int> ri(na * na), ci(na * na);
    for (size_t i = 0, q = 0; i < na; ++i) for (size_t j = 0; j < na; ++j, ++q) { ri[q] = (int)i; ci[q] = (int)j; }
    rc = ordm_set_active_space_sparse(c, di.ncore, di.nact, na * na, ri.data(), ci.data(), (const double*)d1a->data, na * na, ri.data(), ci.data(),
                                      (const double*)d1b->data, d2->count(), ii.data(), jj.data(), kk.data(), ll.data(), (const double*)d2->data, /*upper_packed=*/0);
  }
  if (rc) return rc;
  if (info) *info = di;
  return ORDM_OK;
}

// ---- outputs --------------------------------------------------------------------------------
int ordm_get_vector(ordm_ctx* c, int which, double* host_out, size_t npts) {
  CHECK_CTX(c);
  if (which < 0 || which >= ORDM_VEC_COUNT || (npts && !host_out)) return fail(c, ORDM_ERR_INVALID, "bad vector request");
  if (npts != c->g.npts) return fail(c, ORDM_ERR_INVALID, "vector length %zu != grid size %zu", npts, c->g.npts);
  CU(c, cudaSetDevice(c->device));
  const bool needs_xc = which == ORDM_VEC_R || (which >= ORDM_VEC_TR_RHOA && which <= ORDM_VEC_TR_SIGMA_BB);
  if (needs_xc && !c->xc_done) {
    int rc = staged_xc(c, (c->opts & ORDM_OPT_FULLY_TRANSLATED) != 0);
    if (rc) return rc;
  }
  const double* src = c->g.vec[which];
  // a vector whose producing stage is not current (inputs changed since, or never run) is an error, not old data
  const bool from_rho = which <= ORDM_VEC_SIGMA_BB;
  const bool is_pig = which >= ORDM_VEC_PI_X && which <= ORDM_VEC_PI_Z;
  if (src && ((from_rho && !c->rho_done) || (which == ORDM_VEC_PI && !c->pi_done) || (is_pig && !c->pig_done)))
    return fail(c, ORDM_ERR_STATE, "vector %d is stale: run %s after changing the inputs", which, from_rho ? "build_rho" : "build_pi");
  if (is_pig && !c->pig_done) src = nullptr;  // never built
  if (!src) return fail(c, ORDM_ERR_STATE, "vector %d is not materialised (enable ORDM_OPT_DIAGNOSTICS%s before the build_* calls)", which,
                        which >= ORDM_VEC_PI_X && which <= ORDM_VEC_PI_Z ? " and ORDM_OPT_PI_GRADIENT" : (which >= ORDM_VEC_RHOA_X && which <= ORDM_VEC_SIGMA_BB) || which >= ORDM_VEC_TR_SIGMA_AA ? "; gradients need phi_x/y/z" : "");
  if (npts) CU(c, cudaMemcpyAsync(host_out, src, npts * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return ORDM_OK;
}

int ordm_get_orbitals(ordm_ctx* c, int comp, double* host_out, size_t ld_host) {
  CHECK_CTX(c);
  if (comp < 0 || comp > 3 || !host_out || ld_host < c->g.npts) return fail(c, ORDM_ERR_INVALID, "bad orbital request");
  const double* src = comp == 0 ? c->g.phi : comp == 1 ? c->g.phix : comp == 2 ? c->g.phiy : c->g.phiz;
  if (!src) return fail(c, ORDM_ERR_STATE, "orbital component %d is not set", comp);
  CU(c, cudaSetDevice(c->device));
  if (c->g.npts)
    CU(c, cudaMemcpy2DAsync(host_out, ld_host * sizeof(double), src, c->g.ld * sizeof(double), c->g.npts * sizeof(double), (size_t)c->g.n, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return ORDM_OK;
}

void* ordm_device_vector(ordm_ctx* c, int which) {
  if (!c || which < 0 || which >= ORDM_VEC_COUNT) return nullptr;
  return c->g.vec[which];
}

// ---- synthetic inputs -------------------------------------------------------------------------
int ordm_synth_fill_device(ordm_ctx* c, uint64_t seed, size_t npts_total, size_t p0, size_t m, int n, int gga) {
  CHECK_CTX(c);
  if (n <= 0 || p0 + m > npts_total) return fail(c, ORDM_ERR_INVALID, "bad synthetic shard [%zu,%zu) of %zu", p0, p0 + m, npts_total);
  CU(c, cudaSetDevice(c->device));
  int rc = alloc_w(c, m);
  if (rc) return rc;
  if ((rc = alloc_grid(c, m, n, gga != 0))) return rc;
  if (m) {
    const uint64_t env_key = ordm_stream_key(seed, ORDM_STREAM_ENV);
    dim3 grid((unsigned)((m + 255) / 256), (unsigned)std::min(n, 64));
    for (int k = 0; k < (gga ? 4 : 1); ++k) {
      k_synth_phi<<<grid, 256, 0, c->stream>>>(c->d_phi[k], c->g.ld, m, p0, npts_total, n, ordm_stream_key(seed, ORDM_STREAM_PHI + k), env_key);
      CU(c, cudaGetLastError());
    }
    k_synth_w<<<(unsigned)((m + 255) / 256), 256, 0, c->stream>>>(c->d_w, m, p0, 1.0 / (double)npts_total, ordm_stream_key(seed, ORDM_STREAM_W));
    CU(c, cudaGetLastError());
  }
  CU(c, cudaStreamSynchronize(c->stream));
  return ORDM_OK;
}

int ordm_synth_fill_host(uint64_t seed, size_t npts_total, size_t p0, size_t m, int n, double* w, double* phi, double* px, double* py, double* pz) {
  if (p0 + m > npts_total || n <= 0) return fail(nullptr, ORDM_ERR_INVALID, "bad synthetic shard");
  const uint64_t env_key = ordm_stream_key(seed, ORDM_STREAM_ENV), w_key = ordm_stream_key(seed, ORDM_STREAM_W);
  const double inv = 1.0 / (double)npts_total;
  double* dst[4] = {phi, px, py, pz};
  if (w) for (size_t p = 0; p < m; ++p) w[p] = ordm_synth_w(w_key, inv, p0 + p);
  for (int k = 0; k < 4; ++k) {
    if (!dst[k]) continue;
    const uint64_t key = ordm_stream_key(seed, ORDM_STREAM_PHI + k);
    for (size_t p = 0; p < m; ++p) {
      const double env = ordm_env_key(env_key, p0 + p);
      for (int i = 0; i < n; ++i) dst[k][(size_t)i * m + p] = ordm_synth_phi(key, env, npts_total, p0 + p, (uint64_t)i);
    }
  }
  return ORDM_OK;
}

int ordm_synth_rdms(uint64_t seed, int nact, int na, int nb, int ndet, double* A1a, double* A1b, double* D2act) {
  if (nact <= 0 || na < 0 || nb < 0 || na > nact || nb > nact || ndet <= 0 || !A1a || !A1b)
    return fail(nullptr, ORDM_ERR_INVALID, "bad synthetic RDM request");
  ordm::synth_rdms(seed, nact, na, nb, ndet, A1a, A1b, D2act);
  return ORDM_OK;
}

// ---- multi-GPU -----------------------------------------------------------------------------------
int ordm_comm_unique_id(char id_out[128]) {
  std::string why;
  if (!load_nccl(why)) return fail(nullptr, ORDM_ERR_NCCL, "%s", why.c_str());
  int nr = g_nccl.GetUniqueId(id_out);
  if (nr != 0) return fail(nullptr, ORDM_ERR_NCCL, "ncclGetUniqueId failed (%d)", nr);
  return ORDM_OK;
}

int ordm_comm_attach(ordm_ctx* c, int nranks, int rank, const char id[128]) {
  CHECK_CTX(c);
  if (nranks < 1 || rank < 0 || rank >= nranks || !id) return fail(c, ORDM_ERR_INVALID, "bad communicator geometry");
  std::string why;
  if (!load_nccl(why)) return fail(c, ORDM_ERR_NCCL, "%s", why.c_str());
  CU(c, cudaSetDevice(c->device));
  if (c->comm) ordm_comm_detach(c);   // re-attach: release the previous communicator and its peer mappings first
  IdBlob blob;
  std::memcpy(blob.b, id, 128);
  int nr = g_nccl.CommInitRank(&c->comm, nranks, blob, rank);
  if (nr != 0) { c->comm = nullptr; return fail(c, ORDM_ERR_NCCL, "ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(nr) : "?"); }
  c->nranks = nranks;
  c->rank = rank;
  return setup_peer_xchg(c);  // NVLink peer-store all-reduce when the buffers can be IPC-mapped, else NCCL stays
}

int ordm_comm_info(const ordm_ctx* c, int* nranks, int* rank, int* peer_allreduce) {
  if (!c) return fail(nullptr, ORDM_ERR_INVALID, "null context");
  if (nranks) *nranks = c->nranks;
  if (rank) *rank = c->rank;
  if (peer_allreduce) *peer_allreduce = c->peer_ok ? 1 : 0;
  return ORDM_OK;
}

int ordm_pi_kernel_info(const ordm_ctx* c, int* int8_slice_pairs, uint64_t* mma_per_tile) {
  if (!c) return fail(nullptr, ORDM_ERR_INVALID, "null context");
  if (!c->have_rdm2) return fail(const_cast<ordm_ctx*>(c), ORDM_ERR_STATE, "ordm_pi_kernel_info: no 2-RDM set");
  const bool i8 = c->k2_i8 && !want_pig(c, c->g.gga);
  const int pairs = c->k2_i8_mode == 1 ? 26 : 21;
  if (int8_slice_pairs) *int8_slice_pairs = i8 ? pairs : 0;
  if (mma_per_tile) {
    if (i8) {   // per slice pair: every 32-pair K chunk feeds the right column half, the chunks below it the left half too;
                // the CTA-pair kernel issues one full-width MMA per slice pair and chunk (for a 2 x 128-point tile)
      const int kp = (ordm::pair_count(c->nact) + 31) / 32 * 32;
      *mma_per_tile = c->k2_pair ? (uint64_t)pairs * (uint64_t)(kp / 32) : (uint64_t)pairs * (uint64_t)(kp / 32 + (kp / 2 + 31) / 32);
    } else {
      *mma_per_tile = c->k2_dmma_per_mtile;
    }
  }
  return ORDM_OK;
}

int ordm_comm_detach(ordm_ctx* c) {
  CHECK_CTX(c);
  release_peer_xchg(c);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  c->comm = nullptr;
  c->nranks = 1;
  c->rank = 0;
  return ORDM_OK;
}

void ordm_shard_range(size_t npts_total, int nranks, int rank, size_t* p0, size_t* count) {
  if (nranks < 1) nranks = 1;
  if (rank < 0) rank = 0;
  if (rank >= nranks) rank = nranks - 1;
  // balanced in units of 64 points (the K2 tile / TMA alignment); the last shard takes the ragged tail
  const size_t units = (npts_total + PT_ALIGN - 1) / PT_ALIGN;
  const size_t b = std::min(npts_total, PT_ALIGN * (units * (size_t)rank / (size_t)nranks));
  const size_t e = std::min(npts_total, PT_ALIGN * (units * (size_t)(rank + 1) / (size_t)nranks));
  if (p0) *p0 = b;
  if (count) *count = e - b;
}

}  // extern "C"
