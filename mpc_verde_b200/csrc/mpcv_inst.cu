// mpcv_inst.cu — kernels of ONE model (compiled once per model with -DMPCV_INST_MODEL=<id> so the
// seven models build in parallel) and their launchers.
//
// Two execution shapes (see mpcv_ipm.cuh):
//   *_thread_kernel : one problem per thread.  Workspace is a structure-of-arrays slab in HBM
//       (warp-blocked: element i of problem b at slab[(b/32)*32*total + i*32 + b%32]) so the 32 problems of a warp read and
//       write 256-byte contiguous runs.  Used for tiny problems (nx <= 4, N <= 20-ish).
//   *_warp_kernel   : one problem per warp, persistent grid (CTAs sized to the SM count), per-warp
//       workspace in shared memory, stage-parallel derivative / line-search evaluation, shuffle
//       reductions, TMA bulk copies (cp.async.bulk + mbarrier) staging the per-stage reference
//       window of the problem into shared memory.  Used for long horizons (N = 40..50).
// No tensor cores on purpose: the work is many tiny sequential FP64 factorisations.
#include "mpcv_host.h"

using namespace mpcv;

#ifndef MPCV_INST_MODEL
#error "compile with -DMPCV_INST_MODEL=<model id>"
#endif
#include "mpcv_model_select.h"

// ---------------------------------------------------------------------------------------
// TMA bulk-copy stager for shared-memory workspaces (warp layout)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

struct BulkStager {
  unsigned long long* mbar;   // one mbarrier per warp, in shared memory
  unsigned* phase;            // per-warp phase bit (register copy is re-read from smem each call)

  // Copy `count` doubles from global `src` into ws[off + shift ...]; shift in {0,1} is chosen so
  // that source and destination share the same 16-byte phase; the 16-byte-aligned body moves
  // with one cp.async.bulk (UBLKCP), the at most two ragged elements with plain loads.
  __device__ __forceinline__ int load(const WsDense& ws, int off, const double* src, int count,
                                      const Grp<32>& g) const {
    double* dst0 = ws.base + off;
    const int src_odd = (int)((reinterpret_cast<uintptr_t>(src) >> 3) & 1);
    const int dst_odd = (int)((smem_u32(dst0) >> 3) & 1);
    const int shift = src_odd ^ dst_odd;
    double* dst = dst0 + shift;
    const int head = src_odd;                          // elements before the first 16-byte boundary
    const int body = ((count - head) / 2) * 2;         // even number of doubles
    const int tail = count - head - body;
    if (g.lane < head) dst[g.lane] = src[g.lane];
    if (g.lane < tail) dst[head + body + g.lane] = src[head + body + g.lane];
    if (body > 0) {
      const unsigned bytes = (unsigned)body * 8u;
      const unsigned mb = smem_u32(mbar);
      unsigned ph = *phase;
      __syncwarp();
      if (g.lane == 0) {
        // order earlier generic-proxy accesses to the destination before the async-proxy write
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                smem_u32(dst + head)),
            "l"(src + head), "r"(bytes), "r"(mb)
            : "memory");
      }
      unsigned done = 0;
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(mb), "r"(ph)
            : "memory");
      }
      __syncwarp();
      if (g.lane == 0) *phase = ph ^ 1u;
      __syncwarp();
    }
    return shift;
  }
};

// ---------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <class Model, bool SINGLE>
__global__ void __launch_bounds__(128)
solve_thread_kernel(const Params P, const Layout L, const SolveIO io, double* slab, long B) {
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const long long t0 = io.ns ? globaltimer_ns() : 0;
  solve_problem<Model, SINGLE, 1, WsStrided>(P, L, WsStrided::of(slab, L.total, b), Grp<1>(0), io, b);
  if (io.ns) io.ns[b] = globaltimer_ns() - t0;
}

template <class Model, bool SINGLE>
__global__ void __launch_bounds__(128)
loop_thread_kernel(const Params P, const Layout L, const LoopIO io, double* slab, long B) {
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  closed_loop_problem<Model, SINGLE, 1, WsStrided>(P, L, WsStrided::of(slab, L.total, b), Grp<1>(0), io, b);
}

constexpr int kWarpKernelMaxWarps = 8;

template <class Model>
__global__ void __launch_bounds__(kWarpKernelMaxWarps * 32)
solve_warp_kernel(const Params P, const Layout L, const SolveIO io, long B, int ws_doubles) {
  extern __shared__ __align__(16) double smem[];
  __shared__ unsigned long long mbar[kWarpKernelMaxWarps];
  __shared__ unsigned phase[kWarpKernelMaxWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[warp])) : "memory");
    phase[warp] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const Grp<32> g(lane);
  const BulkStager stager{&mbar[warp], &phase[warp]};
  const WsDense ws{smem + (long)warp * ws_doubles};
  for (long b = (long)blockIdx.x * wpb + warp; b < B; b += (long)gridDim.x * wpb) {
    const long long t0 = io.ns ? globaltimer_ns() : 0;
    solve_problem<Model, false, 32, WsDense, BulkStager>(P, L, ws, g, io, b, stager);
    if (io.ns && lane == 0) io.ns[b] = globaltimer_ns() - t0;
    __syncwarp();
  }
}

template <class Model>
__global__ void __launch_bounds__(kWarpKernelMaxWarps * 32)
loop_warp_kernel(const Params P, const Layout L, const LoopIO io, long B, int ws_doubles) {
  extern __shared__ __align__(16) double smem[];
  __shared__ unsigned long long mbar[kWarpKernelMaxWarps];
  __shared__ unsigned phase[kWarpKernelMaxWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[warp])) : "memory");
    phase[warp] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const Grp<32> g(lane);
  const BulkStager stager{&mbar[warp], &phase[warp]};
  const WsDense ws{smem + (long)warp * ws_doubles};
  for (long b = (long)blockIdx.x * wpb + warp; b < B; b += (long)gridDim.x * wpb) {
    closed_loop_problem<Model, false, 32, WsDense, BulkStager>(P, L, ws, g, io, b, stager);
    __syncwarp();
  }
}

// shooting rollout of a given control sequence (ff(U,P) of single_shooting_v1.py:95, F loop of MS)
template <class Model>
__global__ void rollout_kernel(const Params P, const double* p, const double* U, double* X, double* q, long B) {
  constexpr int NX = Model::NX, NU = Model::NU, NH = NX + Model::NPG;
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int N = P.N, np = NH + N * Model::NPS;
  const double* pb = p + b * np;
  double x[NX], xn[NX], u[NU], qa = 0.0, qk;
#pragma unroll
  for (int i = 0; i < NX; ++i) { x[i] = pb[i]; X[b * NX * (N + 1) + i] = x[i]; }
  for (int k = 0; k < N; ++k) {
#pragma unroll
    for (int i = 0; i < NU; ++i) u[i] = U[b * NU * N + k * NU + i];
    if (Model::HAS_UPREV && P.ntu > 0 && k >= P.ntu) u[0] = x[NX - 1];
    Model::val(P, x, u, pb + NX, pb + NH + k * Model::NPS, xn, &qk);
    qa += qk;
#pragma unroll
    for (int i = 0; i < NX; ++i) { x[i] = xn[i]; X[b * NX * (N + 1) + (k + 1) * NX + i] = xn[i]; }
  }
  if (q) q[b] = qa;
}

template <class Model>
__global__ void stage_derivs_kernel(const Params P, const double* z, const double* pstage, const double* lam,
                                    double* xn, double* A, double* Bm, double* q, double* grad, double* H, long B) {
  constexpr int NX = Model::NX, NU = Model::NU, NZ = NX + NU, NW = NZ * (NZ + 1) / 2;
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double* pp = pstage + b * (Model::NPG + Model::NPS);
  double x[NX], u[NU], l[NX], xo[NX], Ao[NX * NX], Bo[NX * NU], qo, go[NZ], W[NW];
#pragma unroll
  for (int i = 0; i < NX; ++i) { x[i] = z[b * NZ + i]; l[i] = lam[b * NX + i]; }
#pragma unroll
  for (int i = 0; i < NU; ++i) u[i] = z[b * NZ + NX + i];
  Model::der(P, x, u, pp, pp + Model::NPG, l, 1.0, true, xo, Ao, Bo, &qo, go, W);
#pragma unroll
  for (int i = 0; i < NX; ++i) xn[b * NX + i] = xo[i];
#pragma unroll
  for (int i = 0; i < NX * NX; ++i) A[b * NX * NX + i] = Ao[i];
#pragma unroll
  for (int i = 0; i < NX * NU; ++i) Bm[b * NX * NU + i] = Bo[i];
  q[b] = qo;
#pragma unroll
  for (int i = 0; i < NZ; ++i) grad[b * NZ + i] = go[i];
#pragma unroll
  for (int i = 0; i < NZ; ++i) {
#pragma unroll
    for (int j = 0; j < NZ; ++j) H[b * NZ * NZ + i * NZ + j] = i >= j ? W[tri(i, j)] : W[tri(j, i)];
  }
}

template <class Model>
static void fill_dims(mpcv_handle* h) {
  h->nx = Model::NX; h->nu = Model::NU; h->npg = Model::NPG; h->nps = Model::NPS;
  h->has_uprev = Model::HAS_UPREV;
  const int N = h->spec.N;
  h->n_var = h->single ? Model::NU * N : Model::NX * (N + 1) + Model::NU * N;
  h->n_g = Model::NX * (N + 1);
  h->n_p = Model::NX + Model::NPG + N * Model::NPS;
  h->L = h->single ? make_layout<Model, true>(N) : make_layout<Model, false>(N);
}

template <class Model>
static int dims_t(const mpcv_spec* s, int32_t* nx, int32_t* nu, int32_t* n_var, int32_t* n_g, int32_t* n_p,
                  int32_t* npg, int32_t* nps) {
  const bool single = s->shooting == MPCV_SHOOTING_SINGLE;
  if (nx) *nx = Model::NX;
  if (nu) *nu = Model::NU;
  if (n_var) *n_var = single ? Model::NU * s->N : Model::NX * (s->N + 1) + Model::NU * s->N;
  if (n_g) *n_g = Model::NX * (s->N + 1);
  if (n_p) *n_p = Model::NX + Model::NPG + s->N * Model::NPS;
  if (npg) *npg = Model::NPG;
  if (nps) *nps = Model::NPS;
  return 0;
}

static int ensure_slab(mpcv_handle* h, long B) {
  const long stride = (B + 31) / 32 * 32;
  const size_t need = (size_t)stride * h->L.total;
  if (need > h->slab_doubles) {
    if (h->slab) cudaFree(h->slab);
    h->slab = nullptr;
    h->slab_doubles = 0;
    CUDA_OK(cudaMalloc(&h->slab, need * sizeof(double)));
    h->slab_doubles = need;
  }
  h->slab_stride = stride;
  return 0;
}

// warps per CTA and dynamic shared memory of the warp layout
static int warp_config(const mpcv_handle* h, int* wpb, size_t* smem) {
  const size_t per_warp = (size_t)((h->L.total + 1) / 2 * 2) * sizeof(double);
  int w = (int)((h->max_smem_optin - 1024) / per_warp);
  if (w < 1) return mpcv_set_error(-ENOMEM, "problem workspace exceeds shared memory; use MPCV_LAYOUT_THREAD");
  if (w > kWarpKernelMaxWarps) w = kWarpKernelMaxWarps;
  *wpb = w;
  *smem = per_warp * w;
  return 0;
}

template <class Model>
static int launch_solve(mpcv_handle* h, const SolveIO& io, long B, cudaStream_t st) {
  if (B <= 0) return 0;
  if ((h->layout == MPCV_LAYOUT_PHASED || h->layout == MPCV_LAYOUT_RESIDENT) && !h->single)
    return mpcv_phase_vtable_of(Model::MODEL_ID)->solve(h, io, B, st);
  if (h->layout == MPCV_LAYOUT_WARP && !h->single) {
    int wpb; size_t smem;
    if (int rc = warp_config(h, &wpb, &smem)) return rc;
    auto kern = solve_warp_kernel<Model>;
    CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, wpb * 32, smem));
    if (per_sm < 1) per_sm = 1;
    long grid = (long)h->sm_count * per_sm;
    const long need = (B + wpb - 1) / wpb;
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, wpb * 32, smem, st>>>(h->P, h->L, io, B, (int)(smem / sizeof(double) / wpb));
  } else {
    if (int rc = ensure_slab(h, B)) return rc;
    const int threads = 128;
    const unsigned grid = (unsigned)((B + threads - 1) / threads);
    if (h->single) solve_thread_kernel<Model, true><<<grid, threads, 0, st>>>(h->P, h->L, io, h->slab, B);
    else solve_thread_kernel<Model, false><<<grid, threads, 0, st>>>(h->P, h->L, io, h->slab, B);
  }
  h->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

template <class Model>
static int launch_loop(mpcv_handle* h, const LoopIO& io, long B, cudaStream_t st) {
  if (B <= 0) return 0;
  // phased layout: per-step prepare -> solve graph -> apply (mpcv_phase_inst.cu); otherwise the closed
  // loop runs as one kernel per call (warp or thread layout)
  if ((h->layout == MPCV_LAYOUT_PHASED || h->layout == MPCV_LAYOUT_RESIDENT) && !h->single)
    return mpcv_phase_vtable_of(Model::MODEL_ID)->loop(h, io, B, st);
  const bool warp = h->layout == MPCV_LAYOUT_WARP;
  if (warp && !h->single) {
    int wpb; size_t smem;
    if (int rc = warp_config(h, &wpb, &smem)) return rc;
    auto kern = loop_warp_kernel<Model>;
    CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, wpb * 32, smem));
    if (per_sm < 1) per_sm = 1;
    long grid = (long)h->sm_count * per_sm;
    const long need = (B + wpb - 1) / wpb;
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, wpb * 32, smem, st>>>(h->P, h->L, io, B, (int)(smem / sizeof(double) / wpb));
  } else {
    if (int rc = ensure_slab(h, B)) return rc;
    const int threads = 128;
    const unsigned grid = (unsigned)((B + threads - 1) / threads);
    if (h->single) loop_thread_kernel<Model, true><<<grid, threads, 0, st>>>(h->P, h->L, io, h->slab, B);
    else loop_thread_kernel<Model, false><<<grid, threads, 0, st>>>(h->P, h->L, io, h->slab, B);
  }
  h->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

template <class Model>
static int launch_rollout(mpcv_handle* h, const double* p, const double* U, double* X, double* q, long B,
                          cudaStream_t st) {
  if (B <= 0) return 0;
  rollout_kernel<Model><<<(unsigned)((B + 127) / 128), 128, 0, st>>>(h->P, p, U, X, q, B);
  h->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

template <class Model>
static int launch_derivs(mpcv_handle* h, const double* z, const double* pstage, const double* lam, double* xn,
                         double* A, double* Bm, double* q, double* grad, double* H, long B, cudaStream_t st) {
  if (B <= 0) return 0;
  stage_derivs_kernel<Model><<<(unsigned)((B + 127) / 128), 128, 0, st>>>(h->P, z, pstage, lam, xn, A, Bm, q, grad, H, B);
  h->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}


#define MPCV_CAT2(a, b) a##b
#define MPCV_CAT(a, b) MPCV_CAT2(a, b)
extern const mpcv_model_vtable MPCV_CAT(mpcv_model_vtable_, MPCV_INST_MODEL) = {
    dims_t<ModelT>, fill_dims<ModelT>, launch_solve<ModelT>, launch_loop<ModelT>, launch_rollout<ModelT>,
    launch_derivs<ModelT>};
