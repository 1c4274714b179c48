// C-ABI implementation (include/ls_b200.h): context, device memory, kernel launches.
// There is no CPU fallback anywhere in this file: every entry point needs a live CUDA device.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ls_b200.h"
#include "ls_kernels.cuh"

using namespace ls;

#define LS_VERSION 100

struct ls_ctx {
  int device = 0;
  int sm_count = 0;
  int icp_ctas = 0;  // co-resident CTAs for the cooperative ICP kernel
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
  std::string err;
  uint64_t launches = 0;
  // capacities
  int n_cap = 0, m_cap = 0, cells_cap = 0, tab_cap = 0, hist_cap = 0;
  // device buffers
  BuildArrays A{};
  BuildState* bs = nullptr;
  float4 *reading = nullptr, *rd = nullptr;  // raw reading (one-shot path), pre-transformed reading
  float4 *ref_stage = nullptr, *ref_nrm_stage = nullptr;  // one-shot reference staging (scan-frame)
  float* nrm_raw = nullptr;                  // raw normals staging
  size_t nrm_raw_cap = 0;
  int* pos = nullptr;
  float* d2 = nullptr;
  int* ids = nullptr;
  IcpWork* work = nullptr;
  IcpProblem* prob = nullptr;
  float* T_hist = nullptr;
  unsigned long long* phase_ns = nullptr;  // debug (LS_PHASE_TIMING=1)
  float* T0_dev = nullptr;
  // pinned host mirror for small results
  IcpWork* h_work = nullptr;  // only the tail (results) is read
  Grid* h_grid = nullptr;
};

struct ls_scan_slot {
  float4* pts = nullptr;
  float4* nrm = nullptr;
  int n = 0;
  uint64_t id = 0;
  bool used = false;
};

struct ls_map {
  ls_ctx* ctx = nullptr;
  int capacity = 0, max_pts = 0;
  uint64_t next_id = 1;
  std::vector<ls_scan_slot> slots;
};

namespace {

int fail(ls_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  return code;
}

#define CU(call)                                                                                      \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess)                                                                            \
      return fail(ctx, e_ == cudaErrorMemoryAllocation ? LS_ERR_NOMEM : LS_ERR_CUDA, "%s: %s", #call, \
                  cudaGetErrorString(e_));                                                            \
  } while (0)

#define LAUNCH_CHECK()                                                                          \
  do {                                                                                          \
    ++ctx->launches;                                                                            \
    cudaError_t e_ = cudaGetLastError();                                                        \
    if (e_ != cudaSuccess) return fail(ctx, LS_ERR_CUDA, "kernel launch: %s", cudaGetErrorString(e_)); \
  } while (0)

template <typename T>
int dev_alloc(ls_ctx* ctx, T** p, size_t count) {
  if (*p) cudaFree(*p);
  *p = nullptr;
  CU(cudaMalloc((void**)p, count * sizeof(T)));
  return LS_OK;
}

inline int blocks_for(int n, int threads, int cap) {
  int b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  return b > cap ? cap : b;
}

int ensure_capacity(ls_ctx* ctx, int n, int m, int max_cells, int max_iter) {
  if (n > ctx->n_cap) {
    const int cap = n + n / 8 + 1024;
    int rc;
    if ((rc = dev_alloc(ctx, &ctx->reading, (size_t)cap))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->rd, (size_t)cap))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->pos, (size_t)cap))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->d2, (size_t)cap))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->ids, (size_t)cap))) return rc;
    ctx->n_cap = cap;
  }
  if (m > ctx->m_cap) {
    const int cap = m + m / 8 + 1024;
    int rc;
    if ((rc = dev_alloc(ctx, &ctx->A.sub_pts, (size_t)cap))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->A.sub_nrm, (size_t)cap))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->A.srt_pts, (size_t)cap))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->A.srt_nrm, (size_t)cap))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->A.pkey, (size_t)cap))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->ref_stage, (size_t)cap))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->ref_nrm_stage, (size_t)cap))) return rc;
    // a fine table exists only for a level-0 cell with > leaf_split (>= 16) points; the pool is also capped at
    // ~1.5 GB (cells beyond the pool stay leaves: slower, still exact, flagged in stats.grid_overflow)
    int tcap = cap / 17 + 1024;
    const int tmax = (int)((size_t)1536 * 1024 * 1024 / ((size_t)LS_FB3 * 12));
    if (tcap > tmax) tcap = tmax;
    if ((rc = dev_alloc(ctx, &ctx->A.tab1, (size_t)tcap * LS_FB3))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->A.cnt1, (size_t)tcap * LS_FB3))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->A.tab1_cell, (size_t)tcap))) return rc;
    CU(cudaMemsetAsync(ctx->A.cnt1, 0, (size_t)tcap * LS_FB3 * sizeof(uint32_t), ctx->stream));
    ctx->A.tab_cap = tcap;
    ctx->tab_cap = tcap;
    ctx->m_cap = cap;
  }
  if (max_cells > ctx->cells_cap) {
    int rc;
    if ((rc = dev_alloc(ctx, &ctx->A.top, (size_t)max_cells + 1))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->A.cnt0, (size_t)max_cells + 1))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->A.pyr, (size_t)max_cells / 2 + 4096))) return rc;
    CU(cudaMemsetAsync(ctx->A.cnt0, 0, ((size_t)max_cells + 1) * sizeof(uint32_t), ctx->stream));
    ctx->cells_cap = max_cells;
  }
  if (max_iter > ctx->hist_cap) {
    int rc;
    if ((rc = dev_alloc(ctx, &ctx->T_hist, (size_t)max_iter * 16))) return rc;
    ctx->hist_cap = max_iter;
  }
  return LS_OK;
}

struct Resolved {
  float cell;
  int split, max_cells;
};

Resolved resolve(const ls_icp_params* p) {
  Resolved r;
  r.cell = p->cell_size > 0.f ? p->cell_size : 1.0f;
  r.split = p->leaf_split > 0 ? (p->leaf_split < 16 ? 16 : p->leaf_split) : 32;
  r.max_cells = p->max_cells > 0 ? p->max_cells : (1 << 22);
  if (r.max_cells > (1 << 22)) r.max_cells = 1 << 22;  // tile_sums holds 1024 scan tiles
  if (r.max_cells < 64) r.max_cells = 64;
  return r;
}

int check_params(ls_ctx* ctx, const ls_icp_params* p) {
  if (!p) return fail(ctx, LS_ERR_ARG, "null params");
  if (p->max_iterations < 1 || p->max_iterations > 100000) return fail(ctx, LS_ERR_ARG, "max_iterations out of range");
  if (!(p->trim_ratio > 0.f) || p->trim_ratio > 1.f) return fail(ctx, LS_ERR_ARG, "trim_ratio must be in (0,1]");
  if (p->use_differential && (p->smooth_length < 1 || p->smooth_length > kMaxSmooth))
    return fail(ctx, LS_ERR_ARG, "smooth_length must be in [1,%d]", kMaxSmooth);
  return LS_OK;
}

// Build the spatial hash over the sub-map described by `parts` (device pointers), then pre-transform
// the reading.  Everything is enqueued on ctx->stream; nothing synchronises.
int enqueue_build(ls_ctx* ctx, const Parts& parts, const Resolved& r, const float* T0_host) {
  const int m = parts.offset[parts.n_parts];
  reset_build_kernel<<<1, 32, 0, ctx->stream>>>(ctx->bs);
  LAUNCH_CHECK();
  CU(cudaMemcpyAsync(ctx->T0_dev, T0_host, 16 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  const int pb = blocks_for(m, 256, ctx->sm_count * 8);
  assemble_kernel<<<pb, 256, 0, ctx->stream>>>(parts, ctx->A.sub_pts, ctx->A.sub_nrm, ctx->bs);
  LAUNCH_CHECK();
  setup_kernel<<<1, 32, 0, ctx->stream>>>(ctx->bs, m, r.cell, r.max_cells, r.split, ctx->T0_dev);
  LAUNCH_CHECK();
  count0_kernel<<<pb, 256, 0, ctx->stream>>>(ctx->bs, ctx->A, m);
  LAUNCH_CHECK();
  const int tiles = (r.max_cells + kScanTile - 1) / kScanTile;
  scan_reduce_kernel<<<tiles, kScanThreads, 0, ctx->stream>>>(ctx->bs, ctx->A.cnt0);
  LAUNCH_CHECK();
  scan_apply_kernel<<<tiles, kScanThreads, 0, ctx->stream>>>(ctx->bs, ctx->A);
  LAUNCH_CHECK();
  pyramid1_kernel<<<blocks_for(r.max_cells / 16 + 1, 256, ctx->sm_count * 4), 256, 0, ctx->stream>>>(ctx->bs, ctx->A);
  LAUNCH_CHECK();
  pyramid_up_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->bs, ctx->A);
  LAUNCH_CHECK();
  count1_kernel<<<pb, 256, 0, ctx->stream>>>(ctx->bs, ctx->A, m);
  LAUNCH_CHECK();
  tables_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->bs, ctx->A);
  LAUNCH_CHECK();
  scatter_kernel<<<pb, 256, 0, ctx->stream>>>(ctx->A, m);
  LAUNCH_CHECK();
  return LS_OK;
}

int upload_normals(ls_ctx* ctx, const float* normals, int stride, int n, float4* dst) {
  const size_t need = (size_t)n * (size_t)(stride <= 8 ? stride : 3);
  if (need > ctx->nrm_raw_cap) {
    int rc;
    if ((rc = dev_alloc(ctx, &ctx->nrm_raw, need + 4096))) return rc;
    ctx->nrm_raw_cap = need + 4096;
  }
  int dstride = stride;
  if (stride <= 8) {
    CU(cudaMemcpyAsync(ctx->nrm_raw, normals, need * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  } else {
    CU(cudaMemcpy2DAsync(ctx->nrm_raw, 3 * sizeof(float), normals, (size_t)stride * sizeof(float), 3 * sizeof(float),
                         (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    dstride = 3;
  }
  expand_normals_kernel<<<blocks_for(n, 256, ctx->sm_count * 8), 256, 0, ctx->stream>>>(ctx->nrm_raw, dstride, n, dst);
  LAUNCH_CHECK();
  return LS_OK;
}

// Run the persistent ICP kernel on the already-built map + resident reading; fetch results.
int run_icp(ls_ctx* ctx, const ls_icp_params* prm, const float4* reading_dev, int n, const float T0[16],
            float T_out[16], ls_icp_stats* stats, int32_t* opt_ids, float* opt_d2, float* opt_T_hist, int m) {
  reading_kernel<<<blocks_for(n, 256, ctx->sm_count * 8), 256, 0, ctx->stream>>>(ctx->bs, reading_dev, n, ctx->rd);
  LAUNCH_CHECK();
  CU(cudaEventRecord(ctx->ev1, ctx->stream));
  CU(cudaMemsetAsync(ctx->work, 0, sizeof(IcpWork), ctx->stream));
  IcpProblem hp;
  hp.bs = ctx->bs;
  hp.view.top = ctx->A.top;
  hp.view.tab1 = ctx->A.tab1;
  hp.view.pts = ctx->A.srt_pts;
  hp.view.pyr = ctx->A.pyr;
  hp.nrm = ctx->A.srt_nrm;
  hp.rd = ctx->rd;
  hp.n = n;
  hp.pos = ctx->pos;
  hp.d2 = ctx->d2;
  hp.ids = ctx->ids;
  hp.work = ctx->work;
  hp.T_hist = opt_T_hist ? ctx->T_hist : nullptr;
  hp.want_matches = (opt_ids || opt_d2) ? 1 : 0;
  const bool want_phase = getenv("LS_PHASE_TIMING") != nullptr;
  if (want_phase) {
    if (ctx->phase_ns) cudaFree(ctx->phase_ns);
    CU(cudaMalloc((void**)&ctx->phase_ns, (size_t)prm->max_iterations * 6 * sizeof(unsigned long long)));
    CU(cudaMemsetAsync(ctx->phase_ns, 0, (size_t)prm->max_iterations * 6 * sizeof(unsigned long long), ctx->stream));
  }
  hp.phase_ns = want_phase ? ctx->phase_ns : nullptr;
  static unsigned int* warp_cyc_dev = nullptr;
  const bool want_warp = getenv("LS_WARP_PROFILE") != nullptr;
  if (want_warp && !warp_cyc_dev) cudaMalloc((void**)&warp_cyc_dev, 8192 * sizeof(unsigned int));
  if (want_warp) cudaMemsetAsync(warp_cyc_dev, 0, 8192 * sizeof(unsigned int), ctx->stream);
  hp.warp_cyc = want_warp ? warp_cyc_dev : nullptr;
  std::memcpy(hp.T0, T0, sizeof(hp.T0));
  CU(cudaMemcpyAsync(ctx->prob, &hp, sizeof(hp), cudaMemcpyHostToDevice, ctx->stream));
  IcpParamsDev dp;
  dp.max_iterations = prm->max_iterations;
  dp.trim_ratio = prm->trim_ratio;
  dp.use_differential = prm->use_differential;
  dp.min_diff_rot = prm->min_diff_rot;
  dp.min_diff_trans = prm->min_diff_trans;
  dp.smooth_length = prm->smooth_length;
  int ctas = (n + 31) / 32;
  if (ctas > ctx->icp_ctas) ctas = ctx->icp_ctas;
  if (ctas < 1) ctas = 1;
  const IcpProblem* probs = ctx->prob;
  void* args[] = {(void*)&probs, (void*)&ctas, (void*)&dp};
  CU(cudaLaunchCooperativeKernel((void*)icp_kernel, dim3(ctas), dim3(kIcpThreads), args, 0, ctx->stream));
  ++ctx->launches;
  CU(cudaEventRecord(ctx->ev2, ctx->stream));
  // results: small struct + optional arrays
  CU(cudaMemcpyAsync(ctx->h_work, ctx->work, sizeof(IcpWork), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(ctx->h_grid, &ctx->bs->grid, sizeof(Grid), cudaMemcpyDeviceToHost, ctx->stream));
  if (opt_ids) CU(cudaMemcpyAsync(opt_ids, ctx->ids, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (opt_d2) CU(cudaMemcpyAsync(opt_d2, ctx->d2, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  if (opt_T_hist)
    CU(cudaMemcpyAsync(opt_T_hist, ctx->T_hist, (size_t)prm->max_iterations * 16 * sizeof(float), cudaMemcpyDeviceToHost,
                       ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  const IcpWork& w = *ctx->h_work;
  if (want_phase) {
    std::vector<unsigned long long> ph((size_t)prm->max_iterations * 6);
    cudaMemcpy(ph.data(), ctx->phase_ns, ph.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[ls] phase us per iteration: A(nn) B(sel1) C(sel2) D(sel3+acc) E(solve) | total\n");
    for (int it = 0; it < w.iterations && it < prm->max_iterations; ++it) {
      const unsigned long long* q = &ph[(size_t)it * 6];
      fprintf(stderr, "[ls] it %2d: %8.1f %8.1f %8.1f %8.1f %8.1f | %8.1f\n", it, (q[1] - q[0]) * 1e-3, (q[2] - q[1]) * 1e-3,
              (q[3] - q[2]) * 1e-3, (q[4] - q[3]) * 1e-3, (q[5] - q[4]) * 1e-3, (q[5] - q[0]) * 1e-3);
    }
  }
  if (want_warp) {
    std::vector<unsigned int> wc(8192);
    cudaMemcpy(wc.data(), warp_cyc_dev, wc.size() * sizeof(unsigned int), cudaMemcpyDeviceToHost);
    FILE* f = fopen(getenv("LS_WARP_PROFILE"), "wb");
    if (f) { fwrite(wc.data(), sizeof(unsigned int), wc.size(), f); fclose(f); }
  }
  std::memcpy(T_out, w.T_out, 16 * sizeof(float));
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    stats->iterations = w.iterations;
    stats->converged = w.converged;
    stats->max_iter_reached = w.max_iter_reached;
    stats->last_kept = w.last_kept;
    stats->last_limit = w.last_limit;
    stats->used_ratio = n > 0 ? (float)w.last_kept / (float)n : 0.f;
    float ms = 0.f, bms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev2);
    cudaEventElapsedTime(&bms, ctx->ev0, ctx->ev1);
    stats->device_ms = ms;
    stats->build_ms = bms;
    stats->grid_cells = ctx->h_grid->n_cells0;
    stats->grid_tables = ctx->h_grid->n_tab1;
    stats->grid_overflow = ctx->h_grid->overflow;
  }
  if (getenv("LS_DEBUG")) {
    fprintf(stderr, "[ls] status %d fail_code %d iters %d kept %d limit %g total %u bins %u %u %u rem %u %u %u\n", w.status,
            w.fail_code, w.iterations, w.last_kept, w.last_limit, w.dbg_total, w.dbg_bin[0], w.dbg_bin[1], w.dbg_bin[2],
            w.dbg_rem[0], w.dbg_rem[1], w.dbg_rem[2]);
    fprintf(stderr, "[ls] A diag %g %g %g %g %g %g  x %g %g %g %g %g %g\n", w.dbg_A[0], w.dbg_A[1], w.dbg_A[2], w.dbg_A[3],
            w.dbg_A[4], w.dbg_A[5], w.dbg_x[0], w.dbg_x[1], w.dbg_x[2], w.dbg_x[3], w.dbg_x[4], w.dbg_x[5]);
  }
  if (w.status != 0) {
    std::memcpy(T_out, T0, 16 * sizeof(float));
    return fail(ctx, LS_ERR_CONVERGENCE, "ICP: no point to minimise / non-finite transformation");
  }
  (void)m;
  return LS_OK;
}

bool is_identity16(const float* T) {
  for (int i = 0; i < 16; ++i)
    if (T[i] != ((i % 5 == 0) ? 1.f : 0.f)) return false;
  return true;
}

const ls_scan_slot* find_slot(const ls_map* map, uint64_t id) {
  for (const auto& s : map->slots)
    if (s.used && s.id == id) return &s;
  return nullptr;
}

int make_parts(ls_ctx* ctx, const ls_map* map, int n_parts, const uint64_t* part_ids, const float* T_parts, Parts* out) {
  if (n_parts < 1 || n_parts > kMaxParts) return fail(ctx, LS_ERR_ARG, "n_parts must be in [1,%d]", kMaxParts);
  Parts& parts = *out;
  std::memset(&parts, 0, sizeof(parts));
  parts.n_parts = n_parts;
  long long off = 0;
  for (int p = 0; p < n_parts; ++p) {
    const ls_scan_slot* s = find_slot(map, part_ids[p]);
    if (!s) return fail(ctx, LS_ERR_STATE, "scan %llu is not resident (evicted or never pushed)",
                        (unsigned long long)part_ids[p]);
    parts.offset[p] = (int)off;
    parts.pts[p] = s->pts;
    parts.nrm[p] = s->nrm;
    std::memcpy(parts.T[p], T_parts + 16 * p, 16 * sizeof(float));
    parts.identity[p] = is_identity16(T_parts + 16 * p) ? 1 : 0;
    off += s->n;
    if (off > 0x3fffffff) return fail(ctx, LS_ERR_ARG, "sub-map too large");
  }
  parts.offset[n_parts] = (int)off;
  return LS_OK;
}

}  // namespace

extern "C" {

int ls_b200_version(void) { return LS_VERSION; }

int ls_b200_init(int device, ls_ctx** out) {
  if (!out) return LS_ERR_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return LS_ERR_CUDA;  // no CPU fallback
  if (device < 0 || device >= count) return LS_ERR_ARG;
  ls_ctx* ctx = new ls_ctx();
  ctx->device = device;
  auto bail = [&](int code) {
    ls_b200_destroy(ctx);
    return code;
  };
  if (cudaSetDevice(device) != cudaSuccess) return bail(LS_ERR_CUDA);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(LS_ERR_CUDA);
  if (!prop.cooperativeLaunch) return bail(LS_ERR_CUDA);
  ctx->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(LS_ERR_CUDA);
  if (cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
      cudaEventCreate(&ctx->ev2) != cudaSuccess)
    return bail(LS_ERR_CUDA);
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, icp_kernel, kIcpThreads, 0) != cudaSuccess || occ < 1)
    return bail(LS_ERR_CUDA);
  ctx->icp_ctas = occ * ctx->sm_count;
  if (cudaMalloc((void**)&ctx->bs, sizeof(BuildState)) != cudaSuccess) return bail(LS_ERR_NOMEM);
  if (cudaMalloc((void**)&ctx->work, sizeof(IcpWork)) != cudaSuccess) return bail(LS_ERR_NOMEM);
  if (cudaMalloc((void**)&ctx->prob, sizeof(IcpProblem)) != cudaSuccess) return bail(LS_ERR_NOMEM);
  if (cudaMalloc((void**)&ctx->T0_dev, 64 * sizeof(float)) != cudaSuccess) return bail(LS_ERR_NOMEM);
  if (cudaMallocHost((void**)&ctx->h_work, sizeof(IcpWork)) != cudaSuccess) return bail(LS_ERR_NOMEM);
  if (cudaMallocHost((void**)&ctx->h_grid, sizeof(Grid)) != cudaSuccess) return bail(LS_ERR_NOMEM);
  *out = ctx;
  return LS_OK;
}

void ls_b200_destroy(ls_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  void* bufs[] = {ctx->A.sub_pts, ctx->A.sub_nrm, ctx->A.srt_pts, ctx->A.srt_nrm, ctx->A.pkey, ctx->A.top, ctx->A.cnt0,
                  ctx->A.tab1, ctx->A.cnt1, ctx->A.tab1_cell, ctx->A.pyr, ctx->bs,
                  ctx->reading, ctx->rd, ctx->ref_stage, ctx->ref_nrm_stage, ctx->nrm_raw, ctx->pos, ctx->d2, ctx->ids,
                  ctx->work, ctx->prob, ctx->T_hist, ctx->T0_dev};
  for (void* b : bufs)
    if (b) cudaFree(b);
  if (ctx->h_work) cudaFreeHost(ctx->h_work);
  if (ctx->h_grid) cudaFreeHost(ctx->h_grid);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->ev2) cudaEventDestroy(ctx->ev2);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* ls_b200_last_error(const ls_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
uint64_t ls_b200_launch_count(const ls_ctx* ctx) { return ctx ? ctx->launches : 0; }

void ls_icp_default_params(ls_icp_params* p) {
  if (!p) return;
  p->max_iterations = 40;
  p->trim_ratio = 0.75f;
  p->use_differential = 1;
  p->min_diff_rot = 0.001f;
  p->min_diff_trans = 0.01f;
  p->smooth_length = 4;
  p->cell_size = 0.f;
  p->leaf_split = 0;
  p->max_cells = 0;
}

int ls_check_rigid(const float T[16]) { return check_rigid(T); }
void ls_correct_rigid(const float T_in[16], float T_out[16]) { correct_rigid(T_in, T_out); }

int ls_icp_register(ls_ctx* ctx, const ls_icp_params* prm, const float* reading4, int n, const float* ref4,
                    const float* ref_normals, int normals_stride, int m, const float T0[16], float T_out[16],
                    ls_icp_stats* stats, int32_t* opt_ids, float* opt_d2, float* opt_T_iter_hist) {
  if (!ctx) return LS_ERR_ARG;
  if (!reading4 || !ref4 || !ref_normals || !T0 || !T_out || n < 0 || m < 0 || normals_stride < 3)
    return fail(ctx, LS_ERR_ARG, "bad argument");
  int rc = check_params(ctx, prm);
  if (rc) return rc;
  std::memcpy(T_out, T0, 16 * sizeof(float));
  if (stats) std::memset(stats, 0, sizeof(*stats));
  if (n == 0 || m == 0) return fail(ctx, LS_ERR_CONVERGENCE, "empty reading or reference");
  CU(cudaSetDevice(ctx->device));
  const Resolved r = resolve(prm);
  if ((rc = ensure_capacity(ctx, n, m, r.max_cells, prm->max_iterations))) return rc;
  CU(cudaEventRecord(ctx->ev0, ctx->stream));
  CU(cudaMemcpyAsync(ctx->reading, reading4, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->ref_stage, ref4, (size_t)m * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = upload_normals(ctx, ref_normals, normals_stride, m, ctx->ref_nrm_stage))) return rc;
  Parts parts;
  std::memset(&parts, 0, sizeof(parts));
  parts.n_parts = 1;
  parts.offset[0] = 0;
  parts.offset[1] = m;
  parts.pts[0] = ctx->ref_stage;
  parts.nrm[0] = ctx->ref_nrm_stage;
  parts.identity[0] = 1;
  if ((rc = enqueue_build(ctx, parts, r, T0))) return rc;
  return run_icp(ctx, prm, ctx->reading, n, T0, T_out, stats, opt_ids, opt_d2, opt_T_iter_hist, m);
}

int ls_nn_query(ls_ctx* ctx, const ls_icp_params* prm, const float* reading4, int n, const float* ref4, int m,
                const float T0[16], int32_t* ids, float* d2) {
  if (!ctx) return LS_ERR_ARG;
  if (!reading4 || !ref4 || !T0 || !ids || !d2 || n < 0 || m < 0) return fail(ctx, LS_ERR_ARG, "bad argument");
  int rc = check_params(ctx, prm);
  if (rc) return rc;
  if (n == 0) return LS_OK;
  if (m == 0) {
    for (int i = 0; i < n; ++i) { ids[i] = -1; d2[i] = INFINITY; }
    return LS_OK;
  }
  CU(cudaSetDevice(ctx->device));
  const Resolved r = resolve(prm);
  if ((rc = ensure_capacity(ctx, n, m, r.max_cells, 1))) return rc;
  CU(cudaMemcpyAsync(ctx->reading, reading4, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->ref_stage, ref4, (size_t)m * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemsetAsync(ctx->ref_nrm_stage, 0, (size_t)m * sizeof(float4), ctx->stream));
  Parts parts;
  std::memset(&parts, 0, sizeof(parts));
  parts.n_parts = 1;
  parts.offset[1] = m;
  parts.pts[0] = ctx->ref_stage;
  parts.nrm[0] = ctx->ref_nrm_stage;
  parts.identity[0] = 1;
  if ((rc = enqueue_build(ctx, parts, r, T0))) return rc;
  reading_kernel<<<blocks_for(n, 256, ctx->sm_count * 8), 256, 0, ctx->stream>>>(ctx->bs, ctx->reading, n, ctx->rd);
  LAUNCH_CHECK();
  GridView v{ctx->A.top, ctx->A.tab1, ctx->A.srt_pts, ctx->A.pyr};
  nn_query_kernel<<<blocks_for(n, 256, ctx->sm_count * 8), 256, 0, ctx->stream>>>(ctx->bs, v, ctx->rd, n, ctx->ids, ctx->d2);
  LAUNCH_CHECK();
  CU(cudaMemcpyAsync(ids, ctx->ids, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(d2, ctx->d2, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return LS_OK;
}

int ls_transform_cloud(ls_ctx* ctx, const float T[16], const float* in4, const float* normals, int normals_stride,
                       int n, float* out4, float* out_normals3) {
  if (!ctx) return LS_ERR_ARG;
  if (!T || !in4 || !out4 || n < 0 || (normals && normals_stride < 3) || (normals && !out_normals3))
    return fail(ctx, LS_ERR_ARG, "bad argument");
  if (n == 0) return LS_OK;
  CU(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = ensure_capacity(ctx, n, n, 64, 1))) return rc;
  CU(cudaMemcpyAsync(ctx->T0_dev + 16, T, 16 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->reading, in4, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
  if (normals && (rc = upload_normals(ctx, normals, normals_stride, n, ctx->ref_nrm_stage))) return rc;
  transform_kernel<<<blocks_for(n, 256, ctx->sm_count * 8), 256, 0, ctx->stream>>>(
      ctx->T0_dev + 16, ctx->reading, normals ? ctx->ref_nrm_stage : nullptr, n, ctx->rd, ctx->A.sub_nrm);
  LAUNCH_CHECK();
  CU(cudaMemcpyAsync(out4, ctx->rd, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
  if (normals) {
    pack_normals_kernel<<<blocks_for(n, 256, ctx->sm_count * 8), 256, 0, ctx->stream>>>(ctx->A.sub_nrm, n,
                                                                                         (float*)ctx->A.srt_nrm);
    LAUNCH_CHECK();
    CU(cudaMemcpyAsync(out_normals3, ctx->A.srt_nrm, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  }
  CU(cudaStreamSynchronize(ctx->stream));
  return LS_OK;
}

// ---- rolling map ------------------------------------------------------------------------------------
int ls_map_create(ls_ctx* ctx, int capacity_scans, int max_pts_per_scan, ls_map** out) {
  if (!ctx || !out) return LS_ERR_ARG;
  *out = nullptr;
  if (capacity_scans < 2 || capacity_scans > 4096 || max_pts_per_scan < 1)
    return fail(ctx, LS_ERR_ARG, "bad map geometry");
  CU(cudaSetDevice(ctx->device));
  ls_map* map = new ls_map();
  map->ctx = ctx;
  map->capacity = capacity_scans;
  map->max_pts = max_pts_per_scan;
  map->slots.resize(capacity_scans);
  for (auto& s : map->slots) {
    if (cudaMalloc((void**)&s.pts, (size_t)max_pts_per_scan * sizeof(float4)) != cudaSuccess ||
        cudaMalloc((void**)&s.nrm, (size_t)max_pts_per_scan * sizeof(float4)) != cudaSuccess) {
      ls_map_destroy(map);
      return fail(ctx, LS_ERR_NOMEM, "map allocation failed");
    }
  }
  *out = map;
  return LS_OK;
}

void ls_map_destroy(ls_map* map) {
  if (!map) return;
  if (map->ctx) {
    cudaSetDevice(map->ctx->device);
    cudaStreamSynchronize(map->ctx->stream);
  }
  for (auto& s : map->slots) {
    if (s.pts) cudaFree(s.pts);
    if (s.nrm) cudaFree(s.nrm);
  }
  delete map;
}

int ls_map_push_scan(ls_map* map, const float* features4, const float* normals, int normals_stride, int n,
                     uint64_t* scan_id) {
  if (!map) return LS_ERR_ARG;
  ls_ctx* ctx = map->ctx;
  if (!features4 || !normals || normals_stride < 3 || n < 0 || n > map->max_pts || !scan_id)
    return fail(ctx, LS_ERR_ARG, "bad argument (n=%d, max=%d)", n, map->max_pts);
  CU(cudaSetDevice(ctx->device));
  const uint64_t id = map->next_id++;
  ls_scan_slot& s = map->slots[id % (uint64_t)map->capacity];
  s.used = false;
  if (n > 0) {
    CU(cudaMemcpyAsync(s.pts, features4, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    int rc = upload_normals(ctx, normals, normals_stride, n, s.nrm);
    if (rc) return rc;
    // the staging buffer for normals is reused by the next upload: wait for the expand kernel
    CU(cudaStreamSynchronize(ctx->stream));
  }
  s.n = n;
  s.id = id;
  s.used = true;
  *scan_id = id;
  return LS_OK;
}

int ls_map_scan_size(const ls_map* map, uint64_t scan_id) {
  if (!map) return LS_ERR_ARG;
  const ls_scan_slot* s = find_slot(map, scan_id);
  return s ? s->n : LS_ERR_STATE;
}

int ls_icp_register_submap(ls_ctx* ctx, const ls_icp_params* prm, const ls_map* map, uint64_t reading_id, int n_parts,
                           const uint64_t* part_ids, const float* T_parts, const float T0[16], float T_out[16],
                           ls_icp_stats* stats, int32_t* opt_ids, float* opt_d2, float* opt_T_iter_hist) {
  if (!ctx) return LS_ERR_ARG;
  if (!map || map->ctx != ctx || !part_ids || !T_parts || !T0 || !T_out) return fail(ctx, LS_ERR_ARG, "bad argument");
  int rc = check_params(ctx, prm);
  if (rc) return rc;
  std::memcpy(T_out, T0, 16 * sizeof(float));
  if (stats) std::memset(stats, 0, sizeof(*stats));
  const ls_scan_slot* rs = find_slot(map, reading_id);
  if (!rs) return fail(ctx, LS_ERR_STATE, "reading scan %llu is not resident", (unsigned long long)reading_id);
  Parts parts;
  if ((rc = make_parts(ctx, map, n_parts, part_ids, T_parts, &parts))) return rc;
  const int n = rs->n, m = parts.offset[n_parts];
  if (n == 0 || m == 0) return fail(ctx, LS_ERR_CONVERGENCE, "empty reading or reference");
  CU(cudaSetDevice(ctx->device));
  const Resolved r = resolve(prm);
  if ((rc = ensure_capacity(ctx, n, m, r.max_cells, prm->max_iterations))) return rc;
  CU(cudaEventRecord(ctx->ev0, ctx->stream));
  if ((rc = enqueue_build(ctx, parts, r, T0))) return rc;
  return run_icp(ctx, prm, rs->pts, n, T0, T_out, stats, opt_ids, opt_d2, opt_T_iter_hist, m);
}

int ls_map_assemble(ls_ctx* ctx, const ls_map* map, int n_parts, const uint64_t* part_ids, const float* T_parts,
                    float* out4, float* out_normals3, int* m_out) {
  if (!ctx) return LS_ERR_ARG;
  if (!map || map->ctx != ctx || !part_ids || !T_parts || !out4 || !m_out) return fail(ctx, LS_ERR_ARG, "bad argument");
  Parts parts;
  int rc;
  if ((rc = make_parts(ctx, map, n_parts, part_ids, T_parts, &parts))) return rc;
  const int m = parts.offset[n_parts];
  *m_out = m;
  if (m == 0) return LS_OK;
  CU(cudaSetDevice(ctx->device));
  if ((rc = ensure_capacity(ctx, 1, m, 64, 1))) return rc;
  reset_build_kernel<<<1, 32, 0, ctx->stream>>>(ctx->bs);
  LAUNCH_CHECK();
  assemble_kernel<<<blocks_for(m, 256, ctx->sm_count * 8), 256, 0, ctx->stream>>>(parts, ctx->A.sub_pts, ctx->A.sub_nrm,
                                                                                   ctx->bs);
  LAUNCH_CHECK();
  CU(cudaMemcpyAsync(out4, ctx->A.sub_pts, (size_t)m * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
  if (out_normals3) {
    pack_normals_kernel<<<blocks_for(m, 256, ctx->sm_count * 8), 256, 0, ctx->stream>>>(ctx->A.sub_nrm, m,
                                                                                         (float*)ctx->A.srt_nrm);
    LAUNCH_CHECK();
    CU(cudaMemcpyAsync(out_normals3, ctx->A.srt_nrm, (size_t)m * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  }
  CU(cudaStreamSynchronize(ctx->stream));
  return LS_OK;
}

}  // extern "C"
