// mtb_engine.cu — CUDA kernels (sm_100a) and the extern "C" layer of include/mytrim_b200.h.
//
// Kernels:
//   transport_kernel      persistent lanes, one cascade per lane at a time (mtb_transport.cuh)
//   trim_one_kernel       single-ion event mode behind mtb_trim_one
//   stopping_kernel       batched MaterialBase::getrstop
// The host part flattens the caller's configuration into device tables, owns all device
// buffers, launches, and reads tallies back.  No torch types, no exceptions across the ABI.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "mtb_tables.h"
#include "mtb_transport.cuh"

using namespace mtb;

namespace
{
thread_local std::string g_last_error;

int
fail(int code, const std::string & msg)
{
  g_last_error = msg;
  return code;
}

#define MTB_CUDA(call)                                                                                   \
  do                                                                                                     \
  {                                                                                                      \
    cudaError_t err__ = (call);                                                                          \
    if (err__ != cudaSuccess)                                                                            \
      return fail(MTB_ECUDA, std::string(#call) + ": " + cudaGetErrorString(err__));                     \
  } while (0)

#ifndef MTB_BLOCK
#define MTB_BLOCK 128
#endif
constexpr int kBlock = MTB_BLOCK;
// resident CTAs per SM the register allocation of each kernel variant is bounded for (tuned on B200)
#ifndef MTB_MIN_BLOCKS_FAST
#define MTB_MIN_BLOCKS_FAST 7
#endif
#ifndef MTB_MIN_BLOCKS_MONO_NOREC
#define MTB_MIN_BLOCKS_MONO_NOREC 8 // 61 registers without spills; 8 CTAs/SM measured 3.0 % faster than MONO at 7 (r02d)
#endif
#ifndef MTB_MIN_BLOCKS_MONO
#define MTB_MIN_BLOCKS_MONO 7
#endif
#ifndef MTB_MIN_BLOCKS_CLUSTERS
#define MTB_MIN_BLOCKS_CLUSTERS 6
#endif
#ifndef MTB_MIN_BLOCKS_GENERIC
#define MTB_MIN_BLOCKS_GENERIC 6
#endif
#ifndef MTB_MIN_BLOCKS_FAST_SHARE
#define MTB_MIN_BLOCKS_FAST_SHARE 6
#endif
#ifndef MTB_MIN_BLOCKS_MONO_SHARE
#define MTB_MIN_BLOCKS_MONO_SHARE 7 // 72 registers: C->W -1..-2 %, Cu->Cu 150 keV -4.7 % against 6 (the FAST + share kernel keeps 6), visit o
#endif
#ifndef MTB_MIN_BLOCKS_CLUSTERS_SHARE
#define MTB_MIN_BLOCKS_CLUSTERS_SHARE 5 // 93 registers, no spills: -4.5..-7 % on the tests/uo2 workload against 6 (80 registers, spills), r02l
#endif
#ifndef MTB_MIN_BLOCKS_LAYERS_SHARE
#define MTB_MIN_BLOCKS_LAYERS_SHARE 6 // C->W vacenergycount: 5 CTAs/SM measured 9 % slower (r02m)
#endif
#ifndef MTB_MIN_BLOCKS_GENERIC_SHARE
#define MTB_MIN_BLOCKS_GENERIC_SHARE 6
#endif
template <class TR>
constexpr int
min_blocks()
{
  return (TR::kF & F_NOREC)                 ? MTB_MIN_BLOCKS_MONO_NOREC
         : (TR::kF & F_MONO) && !TR::kShare ? MTB_MIN_BLOCKS_MONO
         : (TR::kF & F_MONO)                ? MTB_MIN_BLOCKS_MONO_SHARE
         : TR::kF & F_GEOM_ANY ? (TR::kShare ? MTB_MIN_BLOCKS_GENERIC_SHARE : MTB_MIN_BLOCKS_GENERIC)
         : TR::kF & F_CLUSTERS ? (TR::kShare ? MTB_MIN_BLOCKS_CLUSTERS_SHARE : MTB_MIN_BLOCKS_CLUSTERS)
         : TR::kF & F_FOLLOW   ? (TR::kShare ? MTB_MIN_BLOCKS_LAYERS_SHARE : MTB_MIN_BLOCKS_CLUSTERS)
                                            : (TR::kShare ? MTB_MIN_BLOCKS_FAST_SHARE : MTB_MIN_BLOCKS_FAST);
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
struct SmemLayout
{
  size_t elements, materials, ionz, lowstop, pclass, pairm, paire, layer_cum, layer_mat, hist_vac, hist_repl, blk_u64, blk_f64, pool, pool_ctl, total;
};

__host__ __device__ inline size_t
align_up(size_t v, size_t a)
{
  return (v + a - 1) / a * a;
}

__host__ __device__ inline SmemLayout
smem_layout(const LaunchParams & P)
{
  SmemLayout L;
  size_t o = 0;
  L.blk_u64 = o;
  o += CNT_COUNT * sizeof(unsigned long long);
  L.blk_f64 = o;
  o += 2 * sizeof(double);
  L.layer_cum = o;
  o += (size_t)P.n_layers * sizeof(double);
  L.elements = o = align_up(o, 16);
  o += (size_t)P.n_elements * sizeof(DevElement);
  L.materials = o = align_up(o, 16);
  o += (size_t)P.n_materials * sizeof(DevMaterial);
  L.ionz = o = align_up(o, 16);
  o += (MTB_NZ + 1) * sizeof(DevIonZ);
  L.lowstop = o;
  o += (size_t)(MTB_NZ + 1) * (size_t)P.n_zslots * sizeof(LowStop);
  L.pclass = o = align_up(o, 16);
  o += (size_t)P.n_pclass * sizeof(ProjClass);
  L.pairm = o;
  o += (size_t)P.n_pclass * (size_t)P.n_materials * sizeof(PairM);
  L.paire = o;
  o += (size_t)P.n_pclass * (size_t)P.n_tclass * sizeof(PairE);
  L.layer_mat = o;
  o += (size_t)P.n_layers * sizeof(int32_t);
  L.hist_vac = o = align_up(o, 16);
  o += (size_t)P.smem_hist_bins * sizeof(unsigned int);
  L.hist_repl = o;
  o += (size_t)P.smem_hist_bins * sizeof(unsigned int);
  L.pool = o = align_up(o, 16);
  o += MTB_POOL_SLOTS * sizeof(PoolSlot);
  L.pool_ctl = o;
  o += (POOL_CTL_COUNT + 1) * sizeof(unsigned long long); // + the CTA-local primary counter
  L.total = align_up(o, 16);
  return L;
}

template <class T>
__device__ inline void
block_copy(T * dst, const T * src, size_t n)
{
  const uint32_t * s = reinterpret_cast<const uint32_t *>(src);
  uint32_t * d = reinterpret_cast<uint32_t *>(dst);
  const size_t words = n * sizeof(T) / 4;
  for (size_t i = threadIdx.x; i < words; i += blockDim.x)
    d[i] = s[i];
}

__device__ inline BlockCtx
stage_block(const LaunchParams & P, unsigned char * smem)
{
  const SmemLayout L = smem_layout(P);
  BlockCtx S;
  DevElement * el = reinterpret_cast<DevElement *>(smem + L.elements);
  DevMaterial * mat = reinterpret_cast<DevMaterial *>(smem + L.materials);
  DevIonZ * iz = reinterpret_cast<DevIonZ *>(smem + L.ionz);
  LowStop * lw = reinterpret_cast<LowStop *>(smem + L.lowstop);
  ProjClass * pcl = reinterpret_cast<ProjClass *>(smem + L.pclass);
  PairM * prm = reinterpret_cast<PairM *>(smem + L.pairm);
  PairE * pre = reinterpret_cast<PairE *>(smem + L.paire);
  double * lc = reinterpret_cast<double *>(smem + L.layer_cum);
  int32_t * lm = reinterpret_cast<int32_t *>(smem + L.layer_mat);
  unsigned int * hv = reinterpret_cast<unsigned int *>(smem + L.hist_vac);
  unsigned int * hr = reinterpret_cast<unsigned int *>(smem + L.hist_repl);
  unsigned long long * bu = reinterpret_cast<unsigned long long *>(smem + L.blk_u64);
  double * bf = reinterpret_cast<double *>(smem + L.blk_f64);
  block_copy(el, P.elements, (size_t)P.n_elements);
  block_copy(mat, P.materials, (size_t)P.n_materials);
  block_copy(iz, P.ionz, (size_t)MTB_NZ + 1);
  block_copy(lw, P.lowstop, (size_t)(MTB_NZ + 1) * (size_t)P.n_zslots);
  block_copy(pcl, P.pclass, (size_t)P.n_pclass);
  block_copy(prm, P.pairm, (size_t)P.n_pclass * (size_t)P.n_materials);
  block_copy(pre, P.paire, (size_t)P.n_pclass * (size_t)P.n_tclass);
  if (P.n_layers > 0)
  {
    block_copy(lc, P.layer_cum, (size_t)P.n_layers);
    block_copy(lm, P.layer_mat, (size_t)P.n_layers);
  }
  for (int i = threadIdx.x; i < 2 * P.smem_hist_bins; i += blockDim.x)
    hv[i] = 0u; // hist_repl follows hist_vac
  if (threadIdx.x < CNT_COUNT)
    bu[threadIdx.x] = 0ull;
  if (threadIdx.x < 2)
    bf[threadIdx.x] = 0.0;
  PoolSlot * pool = reinterpret_cast<PoolSlot *>(smem + L.pool);
  unsigned long long * pctl = reinterpret_cast<unsigned long long *>(smem + L.pool_ctl);
  if (threadIdx.x < MTB_POOL_SLOTS)
    pool[threadIdx.x].seq = threadIdx.x;
  if (threadIdx.x <= POOL_CTL_COUNT)
    pctl[threadIdx.x] = threadIdx.x == POOL_WORKING ? (unsigned long long)blockDim.x : 0ull;
  __syncthreads();
  S.elements = el;
  S.materials = mat;
  S.ionz = iz;
  S.lowstop = lw;
  S.pclass = pcl;
  S.pairm = prm;
  S.paire = pre;
  S.layer_cum = lc;
  S.layer_mat = lm;
  S.hist_vac = hv;
  S.hist_repl = hr;
  S.blk_u64 = bu;
  S.blk_f64 = bf;
  S.pool = pool;
  S.pool_ctl = pctl;
  return S;
}

__device__ inline void
flush_block(const LaunchParams & P, const BlockCtx & S)
{
  __syncthreads();
  for (int i = threadIdx.x; i < P.smem_hist_bins; i += blockDim.x)
  {
    const unsigned int v = S.hist_vac[i], r = S.hist_repl[i];
    if (v)
      atomicAdd(&P.u64[off_vac(P) + i], (unsigned long long)v);
    if (r)
      atomicAdd(&P.u64[off_repl(P) + i], (unsigned long long)r);
  }
  if (threadIdx.x < CNT_STACKMAX)
  {
    const unsigned long long v = S.blk_u64[threadIdx.x];
    if (v)
      atomicAdd(&P.u64[threadIdx.x], v);
  }
  if (threadIdx.x == CNT_STACKMAX)
    atomicMax(&P.u64[CNT_STACKMAX], S.blk_u64[CNT_STACKMAX]);
  if (threadIdx.x < 2 && S.blk_f64[threadIdx.x] != 0.0)
    atomicAdd(&P.f64[threadIdx.x], S.blk_f64[threadIdx.x]);
}

// Share kernels run launches with few primaries per lane: they trade one CTA/SM of occupancy for
// 96 registers (no spills in the donation/adoption paths).
template <class TR>
__global__ void __launch_bounds__(kBlock, min_blocks<TR>())
transport_kernel(const __grid_constant__ LaunchParams P)
{
  extern __shared__ __align__(16) unsigned char smem[];
  const BlockCtx S = stage_block(P, smem);
  lane_loop<TR>(P, S, blockIdx.x * blockDim.x + threadIdx.x);
  flush_block(P, S);
}

// rows of the registered primary species, by the functions the lanes use for private rows (mtb_transport.cuh)
__global__ void
primary_class_rows_kernel(const LaunchParams P, ProjClass * pclass, PairM * pairm, PairE * paire)
{
  const int pc = P.n_tclass + blockIdx.x * blockDim.x + threadIdx.x;
  if (pc < P.n_pclass)
    primary_class_rows(pc, P.n_materials, P.n_tclass, P.tmin, P.ionz, P.materials, P.elements, P.tclass_elem, pclass, pairm,
                       paire);
}

__global__ void __launch_bounds__(32)
trim_one_kernel(const __grid_constant__ LaunchParams P)
{
  extern __shared__ __align__(16) unsigned char smem[];
  const BlockCtx S = stage_block(P, smem);
  // lane i follows ion i of the batch (mtb_trim_one: one ion; the other lanes only take part in the warp votes)
  lane_loop<TraitsEvents>(P, S, blockIdx.x * blockDim.x + threadIdx.x);
  flush_block(P, S);
}

__global__ void
stopping_kernel(const __grid_constant__ LaunchParams P, int material, size_t n, const int32_t * Z1,
                const double * m1, const double * E, double * out)
{
  extern __shared__ __align__(16) unsigned char smem[];
  const BlockCtx S = stage_block(P, smem);
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
  {
    const ProjClass pr = make_proj_class(S.ionz[Z1[i]], Z1[i], (float)m1[i]);
    out[i] = (double)material_stopping(pr, S.lowstop + Z1[i] * P.n_zslots, S.materials[material], S.elements, (float)E[i],
                                        fsqrt((float)E[i] * pr.inv_km));
  }
}

// FP32 FMA issue-rate probe: the roofline denominator of this path is the FP32
// pipe, and MEASURED_PEAKS.json only carries HBM and tensor numbers.  8 independent FMA chains per
// thread, 2 flops per FMA.
__global__ void __launch_bounds__(256)
fp32_peak_kernel(float * out, int iters, float a, float b)
{
  float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f,
        x7 = x0 + 7.f;
  for (int i = 0; i < iters; ++i)
  {
#pragma unroll
    for (int u = 0; u < 16; ++u)
    {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  const float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 12345.678f)
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <class T>
struct DevBuf
{
  T * p = nullptr;
  size_t n = 0;
  ~DevBuf() { release(); }
  void release()
  {
    if (p)
      cudaFree(p);
    p = nullptr;
    n = 0;
  }
  cudaError_t ensure(size_t count)
  {
    if (count <= n)
      return cudaSuccess;
    release();
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
    if (e == cudaSuccess)
      n = count;
    return e;
  }
  cudaError_t upload(const T * src, size_t count, cudaStream_t s)
  {
    cudaError_t e = ensure(count);
    if (e != cudaSuccess || !count)
      return e;
    return cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, s);
  }
};
} // namespace

struct mtb_handle
{
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int sm_count = 0;
  int smem_optin = 0; // cudaDevAttrMaxSharedMemoryPerBlockOptin
  int bps[VARIANT_COUNT][2] = {}; // resident CTAs per SM: [variant][share]
  Variant variant = VARIANT_GENERIC, variant_custom = VARIANT_GENERIC; // pick_variant(P, false / true)

  HostConfig host; // host copies of the configuration
  std::vector<int> mat_map; // device material of every input material (mtb_tables.h: fold_identical_materials)
  bool dirty = true, have_materials = false;

  // device tables
  DevBuf<DevElement> d_elements;
  DevBuf<DevMaterial> d_materials;
  DevBuf<DevIonZ> d_ionz;
  DevBuf<LowStop> d_lowstop;
  DevBuf<ProjClass> d_pclass;
  DevBuf<PairM> d_pairm;
  DevBuf<PairE> d_paire;
  DevBuf<int32_t> d_tclass_elem;
  DevBuf<float4_t> d_custom_rows;
  DevBuf<uint32_t> d_deferred;
  bool share_enabled = true;
  uint64_t share_below = 4;  // work sharing for launches with fewer primaries per lane than this
  float share_min_E = 300.f; // eV, see suspend_ion(): measured optimum 200-500 eV (profiles/r02_variant_sweeps.md)
  bool deferred_pending = false;
  float extra_ms = 0.f;
  bool fast = false;
  DevBuf<double> d_layer_cum, d_cl_xyzr;
  DevBuf<int32_t> d_layer_mat, d_cl_hash, d_cl_next;
  DevBuf<uint8_t> d_cl_dist;
  DevBuf<float> d_cl_safe;
  // outputs
  DevBuf<unsigned long long> d_u64;
  DevBuf<double> d_f64;
  DevBuf<mtb_record> d_records;
  DevBuf<mtb_ion_log> d_ionlog;
  DevBuf<RangeEntry> d_range;
  DevBuf<StackEntry> d_stacks;
  void * zero_pinned = nullptr; // page-locked block of zeros (see launch_transport)
  DevBuf<mtb_ion> d_primaries;
  DevBuf<mtb_event> d_events;
  DevBuf<uint32_t> d_event_counts;
  DevBuf<uint64_t> d_uids;
  uint64_t n_resident = 0;
  uint64_t last_n = 0;
  bool records_valid = false;

  LaunchParams P;
  size_t u64_size = 0;
  size_t smem_bytes = 0;
  float last_ms = 0.f;
  bool timing_pending = false;
};

namespace
{
int
build_tables(mtb_handle * h)
{
  if (!h->have_materials)
    return fail(MTB_EINVAL, "mtb_set_materials has not been called");
  const mtb_config & c = h->host.cfg;
  LaunchParams & P = h->P;
  HostTables T;
  std::string err;
  if (int rc = build_host_tables(h->host, T, P, err))
    return fail(rc, err);
  h->mat_map = T.mat_map;

  MTB_CUDA(h->d_elements.upload(T.elements.data(), T.elements.size(), h->stream));
  MTB_CUDA(h->d_materials.upload(T.materials.data(), T.materials.size(), h->stream));
  MTB_CUDA(h->d_ionz.upload(T.ionz.data(), T.ionz.size(), h->stream));
  P.elements = h->d_elements.p;
  P.materials = h->d_materials.p;
  MTB_CUDA(h->d_lowstop.upload(T.lowstop.data(), T.lowstop.size(), h->stream));
  P.ionz = h->d_ionz.p;
  P.lowstop = h->d_lowstop.p;
  MTB_CUDA(h->d_pclass.upload(T.pclass.data(), T.pclass.size(), h->stream));
  MTB_CUDA(h->d_pairm.upload(T.pairm.data(), T.pairm.size(), h->stream));
  MTB_CUDA(h->d_paire.upload(T.paire.data(), T.paire.size(), h->stream));
  P.pclass = h->d_pclass.p;
  P.pairm = h->d_pairm.p;
  P.paire = h->d_paire.p;
  MTB_CUDA(h->d_tclass_elem.upload(T.tclass_elem.data(), T.tclass_elem.size(), h->stream));
  P.tclass_elem = h->d_tclass_elem.p;
  if (P.n_pclass > P.n_tclass)
  {
    primary_class_rows_kernel<<<(P.n_pclass - P.n_tclass + 31) / 32, 32, 0, h->stream>>>(P, h->d_pclass.p, h->d_pairm.p, h->d_paire.p);
    MTB_CUDA(cudaGetLastError());
  }
  if (P.n_layers)
  {
    MTB_CUDA(h->d_layer_cum.upload(T.layer_cum.data(), T.layer_cum.size(), h->stream));
    MTB_CUDA(h->d_layer_mat.upload(T.layer_mat.data(), T.layer_mat.size(), h->stream));
    P.layer_cum = h->d_layer_cum.p;
    P.layer_mat = h->d_layer_mat.p;
  }
  if (P.geom_kind == MTB_GEOM_CLUSTERS)
  {
    MTB_CUDA(h->d_cl_hash.upload(T.cl_hash.data(), T.cl_hash.size(), h->stream));
    MTB_CUDA(h->d_cl_next.upload(T.cl_next.data(), T.cl_next.size(), h->stream));
    MTB_CUDA(h->d_cl_xyzr.upload(h->host.cluster_xyzr.data(), h->host.cluster_xyzr.size(), h->stream));
    P.cl_hash = h->d_cl_hash.p;
    P.cl_next = h->d_cl_next.p;
    P.cl_xyzr = h->d_cl_xyzr.p;
    MTB_CUDA(h->d_cl_dist.upload(T.cl_dist.data(), T.cl_dist.size(), h->stream));
    P.cl_dist = h->d_cl_dist.p;
    P.cl_safe = nullptr;
    if (!T.cl_safe.empty())
    {
      MTB_CUDA(h->d_cl_safe.upload(T.cl_safe.data(), T.cl_safe.size(), h->stream));
      P.cl_safe = h->d_cl_safe.p;
    }
  }
  MTB_CUDA(cudaStreamSynchronize(h->stream)); // T goes out of scope

  const size_t nu64 = u64_block_size(P);
  if (nu64 != h->u64_size)
  {
    h->d_u64.release();
    MTB_CUDA(h->d_u64.ensure(nu64));
    MTB_CUDA(cudaMemsetAsync(h->d_u64.p, 0, nu64 * sizeof(unsigned long long), h->stream));
    h->u64_size = nu64;
  }
  if (!h->d_f64.p)
  {
    MTB_CUDA(h->d_f64.ensure(2));
    MTB_CUDA(cudaMemsetAsync(h->d_f64.p, 0, 2 * sizeof(double), h->stream));
  }
  P.u64 = h->d_u64.p;
  P.f64 = h->d_f64.p;
  if (c.tally_mask & MTB_TALLY_IONLOG)
  {
    MTB_CUDA(h->d_ionlog.ensure(P.ionlog_cap));
    P.ionlog = h->d_ionlog.p;
  }
  if (c.tally_mask & MTB_TALLY_RANGE)
  {
    MTB_CUDA(h->d_range.ensure(P.range_cap));
    P.range = h->d_range.p;
  }

  h->smem_bytes = smem_layout(P).total;
  if (h->smem_bytes > 200 * 1024 || (int)h->smem_bytes > h->smem_optin)
    return fail(MTB_EINVAL, "configuration tables do not fit in shared memory");
  if (std::getenv("MYTRIM_B200_NO_MONO")) // test / tuning knob: single-element samples through the FAST variant
    P.mono = 0;
  h->fast = fast_path_ok(P);
  h->variant = pick_variant(P, false);
  h->variant_custom = pick_variant(P, true);
  if (const char * env = std::getenv("MYTRIM_B200_VARIANT")) // test knob: "generic" forces the all-options kernels
    if (!std::strcmp(env, "generic"))
      h->variant = h->variant_custom = VARIANT_GENERIC;
  // The attribute belongs to the kernel, not to the handle: several engines with different table sizes live in one
  // process (a Trim object's batch and single-ion engines, one engine per configuration in bench.py), so it is set to
  // the device limit once and for all; the occupancy query and the launches use the handle's own size.
  const int smem_limit = h->smem_optin;
  int carveout = -1;
  if (const char * env = std::getenv("MYTRIM_B200_CARVEOUT")) // tuning knob: shared-memory share of the L1 array, percent
    carveout = std::atoi(env);
#define MTB_SETUP_KERNEL(TRAITS, V, SH)                                                                                     \
  MTB_CUDA(cudaFuncSetAttribute(transport_kernel<TRAITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit)); \
  if (carveout >= 0)                                                                                                        \
    MTB_CUDA(cudaFuncSetAttribute(transport_kernel<TRAITS>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout));   \
  MTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->bps[V][SH], transport_kernel<TRAITS>, kBlock, h->smem_bytes));  \
  h->bps[V][SH] = std::max(h->bps[V][SH], 1);
  MTB_SETUP_KERNEL(TraitsFast, VARIANT_FAST, 0)
  MTB_SETUP_KERNEL(TraitsFastShare, VARIANT_FAST, 1)
  MTB_SETUP_KERNEL(TraitsMono, VARIANT_MONO, 0)
  MTB_SETUP_KERNEL(TraitsMonoShare, VARIANT_MONO, 1)
  MTB_SETUP_KERNEL(TraitsMonoNoRec, VARIANT_MONO_NOREC, 0)
  h->bps[VARIANT_MONO_NOREC][1] = h->bps[VARIANT_MONO][1]; // small launches share through the MONO twin
  MTB_SETUP_KERNEL(TraitsClusters, VARIANT_CLUSTERS, 0)
  MTB_SETUP_KERNEL(TraitsClustersShare, VARIANT_CLUSTERS, 1)
  MTB_SETUP_KERNEL(TraitsFastPhonon, VARIANT_FAST_PHONON, 0)
  MTB_SETUP_KERNEL(TraitsFastPhononShare, VARIANT_FAST_PHONON, 1)
  MTB_SETUP_KERNEL(TraitsMonoEvac, VARIANT_MONO_EVAC, 0)
  MTB_SETUP_KERNEL(TraitsMonoEvacShare, VARIANT_MONO_EVAC, 1)
  MTB_SETUP_KERNEL(TraitsClustersLog, VARIANT_CLUSTERS_LOG, 0)
  MTB_SETUP_KERNEL(TraitsClustersLogShare, VARIANT_CLUSTERS_LOG, 1)
  MTB_SETUP_KERNEL(TraitsLayersPlain, VARIANT_LAYERS_PLAIN, 0)
  MTB_SETUP_KERNEL(TraitsLayersPlainShare, VARIANT_LAYERS_PLAIN, 1)
  MTB_SETUP_KERNEL(TraitsLayers, VARIANT_LAYERS, 0)
  MTB_SETUP_KERNEL(TraitsLayersShare, VARIANT_LAYERS, 1)
  MTB_SETUP_KERNEL(TraitsGeneric, VARIANT_GENERIC, 0)
  MTB_SETUP_KERNEL(TraitsGenericShare, VARIANT_GENERIC, 1)
#undef MTB_SETUP_KERNEL
  MTB_CUDA(cudaFuncSetAttribute(trim_one_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit));
  MTB_CUDA(cudaFuncSetAttribute(stopping_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit));
  if (const char * cap = std::getenv("MYTRIM_B200_BLOCKS_PER_SM")) // tuning knob: resident CTAs per SM
    for (int v = 0; v < VARIANT_COUNT; ++v)
      h->bps[v][0] = std::max(1, std::min(h->bps[v][0], std::atoi(cap)));
  h->dirty = false;
  return MTB_OK;
}

int
ensure_ready(mtb_handle * h)
{
  if (!h)
    return fail(MTB_EINVAL, "null handle");
  MTB_CUDA(cudaSetDevice(h->device));
  if (h->dirty)
    return build_tables(h);
  return MTB_OK;
}

void
launch_kernel(mtb_handle * h, const LaunchParams & P, unsigned blocks, Variant v, bool share)
{
  switch (v)
  {
    case VARIANT_FAST:
      if (share)
        transport_kernel<TraitsFastShare><<<blocks, kBlock, h->smem_bytes, h->stream>>>(P);
      else
        transport_kernel<TraitsFast><<<blocks, kBlock, h->smem_bytes, h->stream>>>(P);
      break;
    case VARIANT_MONO_NOREC:
      if (share)
        transport_kernel<TraitsMonoShare><<<blocks, kBlock, h->smem_bytes, h->stream>>>(P);
      else
        transport_kernel<TraitsMonoNoRec><<<blocks, kBlock, h->smem_bytes, h->stream>>>(P);
      break;
    case VARIANT_MONO:
      if (share)
        transport_kernel<TraitsMonoShare><<<blocks, kBlock, h->smem_bytes, h->stream>>>(P);
      else
        transport_kernel<TraitsMono><<<blocks, kBlock, h->smem_bytes, h->stream>>>(P);
      break;
    case VARIANT_CLUSTERS:
      if (share)
        transport_kernel<TraitsClustersShare><<<blocks, kBlock, h->smem_bytes, h->stream>>>(P);
      else
        transport_kernel<TraitsClusters><<<blocks, kBlock, h->smem_bytes, h->stream>>>(P);
      break;
    case VARIANT_FAST_PHONON:
      if (share)
        transport_kernel<TraitsFastPhononShare><<<blocks, kBlock, h->smem_bytes, h->stream>>>(P);
      else
        transport_kernel<TraitsFastPhonon><<<blocks, kBlock, h->smem_bytes, h->stream>>>(P);
      break;
    case VARIANT_MONO_EVAC:
      if (share)
        transport_kernel<TraitsMonoEvacShare><<<blocks, kBlock, h->smem_bytes, h->stream>>>(P);
      else
        transport_kernel<TraitsMonoEvac><<<blocks, kBlock, h->smem_bytes, h->stream>>>(P);
      break;
    case VARIANT_CLUSTERS_LOG:
      if (share)
        transport_kernel<TraitsClustersLogShare><<<blocks, kBlock, h->smem_bytes, h->stream>>>(P);
      else
        transport_kernel<TraitsClustersLog><<<blocks, kBlock, h->smem_bytes, h->stream>>>(P);
      break;
    case VARIANT_LAYERS_PLAIN:
      if (share)
        transport_kernel<TraitsLayersPlainShare><<<blocks, kBlock, h->smem_bytes, h->stream>>>(P);
      else
        transport_kernel<TraitsLayersPlain><<<blocks, kBlock, h->smem_bytes, h->stream>>>(P);
      break;
    case VARIANT_LAYERS:
      if (share)
        transport_kernel<TraitsLayersShare><<<blocks, kBlock, h->smem_bytes, h->stream>>>(P);
      else
        transport_kernel<TraitsLayers><<<blocks, kBlock, h->smem_bytes, h->stream>>>(P);
      break;
    default:
      if (share)
        transport_kernel<TraitsGenericShare><<<blocks, kBlock, h->smem_bytes, h->stream>>>(P);
      else
        transport_kernel<TraitsGeneric><<<blocks, kBlock, h->smem_bytes, h->stream>>>(P);
  }
}

// grid of a launch of n primaries with variant v: resident CTAs only; work sharing below share_below
// primaries per lane
unsigned
launch_grid(const mtb_handle * h, Variant v, uint64_t n, bool * share)
{
  const uint64_t max_blocks = (uint64_t)h->sm_count * h->bps[v][0];
  *share = h->share_enabled && n < h->share_below * max_blocks * kBlock;
  return (unsigned)(*share ? (uint64_t)h->sm_count * h->bps[v][1] : std::min<uint64_t>(max_blocks, (n + kBlock - 1) / kBlock));
}

int drain(mtb_handle * h);
constexpr size_t kZeroBlock = 4u << 20;

int
launch_transport(mtb_handle * h, uint64_t n, const mtb_ion * primaries_dev, const mtb_ion * beam, uint64_t seed,
                 uint64_t first_index, bool want_records)
{
  // a previous asynchronous launch may still owe its deferred primaries (their list is re-used below)
  if (h->deferred_pending)
    if (int rc = drain(h))
      return rc;
  LaunchParams & P = h->P;
  P.primaries = primaries_dev;
  if (beam)
    P.beam = *beam;
  P.n_primaries = n;
  P.first_index = first_index;
  P.key0 = (uint32_t)seed;
  P.key1 = (uint32_t)(seed >> 32);
  philox_round_keys(P.key0, P.key1, P.rk);
  P.share_min_E = h->share_min_E;
  P.records = nullptr;
  if (want_records)
  {
    MTB_CUDA(h->d_records.ensure(n));
    P.records = h->d_records.p;
  }
  h->records_valid = want_records;
  h->last_n = n;
  if (!n)
    return MTB_OK;
  // beam mode with a species that has no projectile class cannot defer: take a variant with F_CUSTOM
  Variant v = (primaries_dev || species_known(h->host, P.beam.Z, P.beam.m)) ? h->variant : h->variant_custom;
  if (v == VARIANT_MONO && !want_records && !std::getenv("MYTRIM_B200_NO_NOREC"))
    v = VARIANT_MONO_NOREC; // same arithmetic, record bookkeeping compiled out
  bool share;
  const unsigned blocks = launch_grid(h, v, n, &share);
  if ((uint64_t)blocks * kBlock * MTB_STACK_DEPTH * sizeof(StackEntry) > 0xFFFFFFFFull)
    return fail(MTB_EINVAL, "grid too large for 32-bit stack cursors");
  if (want_records)
  {
    // lanes accumulate into the records: they start from zero.  A launch of a work-sharing kernel zeroes them with the
    // copy engine (from a page-locked block of zeros) instead of a memset KERNEL: when a second engine of the same GPU
    // still runs its persistent kernel, a memset kernel finds no free registers on any SM and this stream — and the host
    // thread behind it — would wait for that kernel's last CTA (measured in apps/mytrim_uo2: launches of two engines
    // never overlapped).  Large launches keep the memset kernel (3 TB/s instead of PCIe).
    const size_t bytes = n * sizeof(mtb_record);
    if (share && bytes <= (256u << 20))
    {
      if (!h->zero_pinned)
      {
        MTB_CUDA(cudaHostAlloc(&h->zero_pinned, kZeroBlock, cudaHostAllocDefault));
        std::memset(h->zero_pinned, 0, kZeroBlock);
      }
      for (size_t off = 0; off < bytes; off += kZeroBlock)
        MTB_CUDA(cudaMemcpyAsync(reinterpret_cast<char *>(h->d_records.p) + off, h->zero_pinned, std::min(kZeroBlock, bytes - off),
                                 cudaMemcpyHostToDevice, h->stream));
    }
    else
      MTB_CUDA(cudaMemsetAsync(h->d_records.p, 0, bytes, h->stream));
  }
  MTB_CUDA(h->d_stacks.ensure((size_t)blocks * kBlock * MTB_STACK_DEPTH));
  P.stacks = h->d_stacks.p;
  MTB_CUDA(h->d_custom_rows.ensure((size_t)blocks * kBlock * (size_t)(2 + P.n_materials + P.n_tclass)));
  P.custom_rows = h->d_custom_rows.p;
  // variants without F_CUSTOM hand class-less primaries to a second launch
  const bool defers = !(variant_features(v) & F_CUSTOM) && primaries_dev != nullptr;
  P.index_list = nullptr;
  P.deferred = nullptr;
  if (defers)
  {
    if (n > 0xFFFFFFFFull)
      return fail(MTB_EINVAL, "more than 2^32 primaries in one launch");
    MTB_CUDA(h->d_deferred.ensure(n));
    P.deferred = h->d_deferred.p;
  }
  MTB_CUDA(cudaMemsetAsync(&P.u64[CNT_NEXT_PRIMARY], 0, sizeof(unsigned long long), h->stream));
  MTB_CUDA(cudaMemsetAsync(&P.u64[CNT_DEFERRED], 0, sizeof(unsigned long long), h->stream));
  MTB_CUDA(cudaEventRecord(h->ev0, h->stream));
  launch_kernel(h, P, blocks, v, share);
  MTB_CUDA(cudaGetLastError());
  MTB_CUDA(cudaEventRecord(h->ev1, h->stream));
  h->timing_pending = true;
  h->deferred_pending = defers;
  h->extra_ms = 0.f;
  return MTB_OK;
}

// Primaries the fast kernel could not take (their species has no projectile class) run through the
// generic kernel, which builds per-lane class rows for them.
int
run_deferred(mtb_handle * h)
{
  h->deferred_pending = false;
  unsigned long long nd = 0;
  MTB_CUDA(cudaMemcpy(&nd, h->d_u64.p + CNT_DEFERRED, sizeof(nd), cudaMemcpyDeviceToHost));
  if (!nd)
    return MTB_OK;
  LaunchParams P = h->P;
  P.index_list = h->d_deferred.p;
  P.deferred = nullptr;
  P.n_primaries = nd;
  bool share;
  const unsigned blocks = launch_grid(h, h->variant_custom, nd, &share);
  if ((uint64_t)blocks * kBlock * MTB_STACK_DEPTH * sizeof(StackEntry) > 0xFFFFFFFFull)
    return fail(MTB_EINVAL, "grid too large for 32-bit stack cursors");
  MTB_CUDA(h->d_stacks.ensure((size_t)blocks * kBlock * MTB_STACK_DEPTH));
  P.stacks = h->d_stacks.p;
  MTB_CUDA(h->d_custom_rows.ensure((size_t)blocks * kBlock * (size_t)(2 + P.n_materials + P.n_tclass)));
  P.custom_rows = h->d_custom_rows.p;
  MTB_CUDA(cudaMemsetAsync(&P.u64[CNT_NEXT_PRIMARY], 0, sizeof(unsigned long long), h->stream));
  cudaEvent_t e0, e1;
  MTB_CUDA(cudaEventCreate(&e0));
  MTB_CUDA(cudaEventCreate(&e1));
  MTB_CUDA(cudaEventRecord(e0, h->stream));
  launch_kernel(h, P, blocks, h->variant_custom, share);
  MTB_CUDA(cudaGetLastError());
  MTB_CUDA(cudaEventRecord(e1, h->stream));
  MTB_CUDA(cudaStreamSynchronize(h->stream));
  MTB_CUDA(cudaEventElapsedTime(&h->extra_ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return MTB_OK;
}

int
sync_and_check(mtb_handle * h)
{
  MTB_CUDA(cudaStreamSynchronize(h->stream));
  if (h->timing_pending)
  {
    MTB_CUDA(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    h->timing_pending = false;
  }
  if (h->deferred_pending)
  {
    if (int rc = run_deferred(h))
      return rc;
    h->last_ms += h->extra_ms;
  }
  unsigned long long err = 0;
  MTB_CUDA(cudaMemcpy(&err, h->d_u64.p + CNT_ERROR, sizeof(err), cudaMemcpyDeviceToHost));
  if (err)
  {
    // the word is cleared so that the handle stays usable; the tallies hold everything that was followed
    MTB_CUDA(cudaMemsetAsync(h->d_u64.p + CNT_ERROR, 0, sizeof(err), h->stream));
    MTB_CUDA(cudaStreamSynchronize(h->stream));
    if (err >> 32)
      return fail(MTB_EINVAL, std::to_string(err >> 32) + " primaries were skipped: Z outside 1..92, mass <= 0, negative or "
                                                          "non-finite energy, or no direction");
    return fail(MTB_ESTACK, "recoil stack overflow in " + std::to_string(err & 0xFFFFFFFFull) + " collision(s)");
  }
  return MTB_OK;
}

// Everything an asynchronous mtb_launch_resident left behind — the launch of the deferred (class-less) primaries and
// the stack-overflow check — has to happen before tallies are read, reduced or reset and before the next launch
// re-uses the deferral list.
// event mode (mtb_trim_one / mtb_trim_many): an ion the tables cannot describe was skipped on the device
int
check_skipped_ions(mtb_handle * h)
{
  unsigned long long err = 0;
  MTB_CUDA(cudaMemcpy(&err, h->d_u64.p + CNT_ERROR, sizeof(err), cudaMemcpyDeviceToHost));
  if (!err)
    return MTB_OK;
  MTB_CUDA(cudaMemset(h->d_u64.p + CNT_ERROR, 0, sizeof(err)));
  return fail(MTB_EINVAL, std::to_string(err >> 32) + " ion(s) were skipped: Z outside 1..92, mass <= 0, negative or non-finite "
                                                      "energy, or no direction");
}

int
drain(mtb_handle * h)
{
  if (h->timing_pending || h->deferred_pending)
    return sync_and_check(h);
  MTB_CUDA(cudaStreamSynchronize(h->stream));
  return MTB_OK;
}
} // namespace

// ---------------------------------------------------------------------------------------------
// extern "C"
// ---------------------------------------------------------------------------------------------
extern "C" {

const char *
mtb_version(void)
{
  return "mytrim_b200 0.1 (sm_100a)";
}

const char *
mtb_last_error(void)
{
  return g_last_error.c_str();
}

int
mtb_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess)
    return 0;
  return n;
}

void
mtb_default_config(mtb_config * cfg)
{
  default_config(cfg);
}

int
mtb_create(const mtb_config * cfg, mtb_handle ** out)
{
  if (!cfg || !out)
    return fail(MTB_EINVAL, "null argument");
  if (!(cfg->length_scale > 0.0) || cfg->potential < 0 || cfg->potential > 2)
    return fail(MTB_EINVAL, "bad configuration");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(MTB_ENODEV, "no CUDA device: the transport engine has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev)
    return fail(MTB_EINVAL, "device ordinal out of range");
  MTB_CUDA(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  MTB_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major < 10)
    return fail(MTB_ENODEV, std::string("device ") + prop.name + " is not sm_100 class");
  mtb_handle * h = new (std::nothrow) mtb_handle();
  if (!h)
    return fail(MTB_ENOMEM, "out of memory");
  h->host.cfg = *cfg;
  h->device = cfg->device;
  if (const char * env = std::getenv("MYTRIM_B200_NO_SHARE"))
    h->share_enabled = env[0] == '0';
  if (const char * env = std::getenv("MYTRIM_B200_SHARE_MIN_E")) // tuning knob: eV
    h->share_min_E = (float)std::atof(env);
  if (const char * env = std::getenv("MYTRIM_B200_SHARE_BELOW")) // tuning knob: primaries per lane
    h->share_below = std::strtoull(env, nullptr, 10);
  h->sm_count = prop.multiProcessorCount;
  h->smem_optin = (int)prop.sharedMemPerBlockOptin;
  cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess)
    e = cudaEventCreate(&h->ev0);
  if (e == cudaSuccess)
    e = cudaEventCreate(&h->ev1);
  if (e != cudaSuccess)
  {
    delete h;
    return fail(MTB_ECUDA, cudaGetErrorString(e));
  }
  *out = h;
  return MTB_OK;
}

int
mtb_destroy(mtb_handle * h)
{
  if (!h)
    return MTB_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  cudaEventDestroy(h->ev0);
  cudaEventDestroy(h->ev1);
  cudaStreamDestroy(h->stream);
  if (h->zero_pinned)
    cudaFreeHost(h->zero_pinned);
  delete h;
  return MTB_OK;
}

int
mtb_get_tables(double * pcoef, double * vfermi, double * lfctr, double * mm1)
{
  for (int z = 0; z < MTB_NZ; ++z)
  {
    if (pcoef)
      std::memcpy(pcoef + 8 * z, builtin_zbl()[z].pcoef, 8 * sizeof(double));
    if (vfermi)
      vfermi[z] = builtin_zbl()[z].vfermi;
    if (lfctr)
      lfctr[z] = builtin_zbl()[z].lfctr;
    if (mm1)
      mm1[z] = builtin_zbl()[z].mm1;
  }
  return MTB_OK;
}

int
mtb_set_tables(mtb_handle * h, const double * pcoef, const double * vfermi, const double * lfctr, const double * mm1)
{
  if (!h)
    return fail(MTB_EINVAL, "null handle");
  for (int z = 0; z < MTB_NZ; ++z)
  {
    if (pcoef)
      std::memcpy(h->host.zbl[z].pcoef, pcoef + 8 * z, 8 * sizeof(double));
    if (vfermi)
      h->host.zbl[z].vfermi = vfermi[z];
    if (lfctr)
      h->host.zbl[z].lfctr = lfctr[z];
    if (mm1)
      h->host.zbl[z].mm1 = mm1[z];
  }
  h->dirty = true;
  return MTB_OK;
}

int
mtb_set_materials(mtb_handle * h, int n_materials, const mtb_material * materials, int n_elements,
                  const mtb_element * elements)
{
  if (!h)
    return fail(MTB_EINVAL, "null handle");
  std::string err;
  if (int rc = check_materials(n_materials, materials, n_elements, elements, err))
    return fail(rc, err);
  h->host.materials.assign(materials, materials + n_materials);
  h->host.elements.assign(elements, elements + n_elements);
  h->have_materials = true;
  h->dirty = true;
  return MTB_OK;
}

int
mtb_set_geometry(mtb_handle * h, const mtb_geometry * g)
{
  if (!h)
    return fail(MTB_EINVAL, "null handle");
  std::string err;
  if (int rc = check_geometry(g, err))
    return fail(rc, err);
  h->host.layer_thickness.clear();
  h->host.cluster_xyzr.clear();
  if (g->kind == MTB_GEOM_LAYERS)
    h->host.layer_thickness.assign(g->layer_thickness, g->layer_thickness + g->n_layers);
  if (g->kind == MTB_GEOM_CLUSTERS && g->n_clusters)
    h->host.cluster_xyzr.assign(g->cluster_xyzr, g->cluster_xyzr + 4 * (size_t)g->n_clusters);
  h->host.geom = *g;
  h->host.geom.layer_thickness = nullptr;
  h->host.geom.cluster_xyzr = nullptr;
  h->dirty = true;
  return MTB_OK;
}

int
mtb_upload_primaries(mtb_handle * h, uint64_t n, const mtb_ion * primaries)
{
  if (!h)
    return fail(MTB_EINVAL, "null handle");
  if (n && !primaries)
    return fail(MTB_EINVAL, "null primaries");
  // species of the primaries get their own rows in the class tables
  if (h->have_materials && register_primary_species(h->host, std::min<uint64_t>(n, 64), primaries))
    h->dirty = true;
  if (int rc = ensure_ready(h))
    return rc;
  MTB_CUDA(h->d_primaries.upload(primaries, n, h->stream));
  h->n_resident = n;
  return MTB_OK;
}

int
mtb_launch_resident(mtb_handle * h, uint64_t seed, uint64_t first_index)
{
  if (int rc = ensure_ready(h))
    return rc;
  return launch_transport(h, h->n_resident, h->d_primaries.p, nullptr, seed, first_index,
                          (h->host.cfg.tally_mask & MTB_TALLY_RECORDS) != 0);
}

int
mtb_synchronize(mtb_handle * h)
{
  if (!h)
    return fail(MTB_EINVAL, "null handle");
  MTB_CUDA(cudaSetDevice(h->device));
  return sync_and_check(h);
}

int
mtb_last_kernel_ms(mtb_handle * h, float * ms)
{
  if (!h || !ms)
    return fail(MTB_EINVAL, "null argument");
  if (h->timing_pending)
    if (int rc = mtb_synchronize(h))
      return rc;
  *ms = h->last_ms;
  return MTB_OK;
}

const char *
mtb_kernel_variant(mtb_handle * h)
{
  if (ensure_ready(h))
    return "";
  return variant_name(h->variant);
}

int
mtb_fetch_records(mtb_handle * h, uint64_t n, mtb_record * records)
{
  if (!h || !records)
    return fail(MTB_EINVAL, "null argument");
  if (!h->records_valid || n > h->last_n)
    return fail(MTB_EINVAL, "no records: enable MTB_TALLY_RECORDS");
  MTB_CUDA(cudaSetDevice(h->device));
  MTB_CUDA(cudaMemcpyAsync(records, h->d_records.p, n * sizeof(mtb_record), cudaMemcpyDeviceToHost, h->stream));
  MTB_CUDA(cudaStreamSynchronize(h->stream));
  return MTB_OK;
}

int
mtb_run(mtb_handle * h, uint64_t n, const mtb_ion * primaries, uint64_t seed, uint64_t first_index,
        mtb_record * records)
{
  if (!h)
    return fail(MTB_EINVAL, "null handle");
  if (n && !primaries)
    return fail(MTB_EINVAL, "null primaries");
  // Page-locked caller memory (cudaHostAlloc / cudaHostRegister, e.g. a pinned torch tensor) is read
  // by the lanes in place over PCIe: each primary is 88 bytes fetched once per cascade, so the
  // transfer hides behind the transport instead of preceding it.  Pageable memory is staged.
  const mtb_ion * dev_view = nullptr;
  if (n)
  {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, primaries) == cudaSuccess)
    {
      if (attr.type == cudaMemoryTypeDevice)
        return fail(MTB_EINVAL, "mtb_run takes host pointers (use mtb_upload_primaries + mtb_launch_resident)");
      if (attr.type == cudaMemoryTypeHost && attr.devicePointer)
        dev_view = static_cast<const mtb_ion *>(attr.devicePointer);
    }
    else
      (void)cudaGetLastError();
  }
  if (dev_view)
  {
    if (h->have_materials && register_primary_species(h->host, std::min<uint64_t>(n, 64), primaries))
      h->dirty = true;
    if (int rc = ensure_ready(h))
      return rc;
  }
  else
  {
    if (int rc = mtb_upload_primaries(h, n, primaries))
      return rc;
    dev_view = h->d_primaries.p;
  }
  if (int rc = launch_transport(h, n, dev_view, nullptr, seed, first_index, records != nullptr))
    return rc;
  if (int rc = sync_and_check(h))
    return rc;
  if (records && n)
    return mtb_fetch_records(h, n, records);
  return MTB_OK;
}

int
mtb_run_beam(mtb_handle * h, uint64_t n, const mtb_ion * ion, uint64_t seed, uint64_t first_index,
             mtb_record * records)
{
  if (!h || !ion)
    return fail(MTB_EINVAL, "null argument");
  if (h->have_materials && register_primary_species(h->host, 1, ion))
    h->dirty = true;
  if (int rc = ensure_ready(h))
    return rc;
  if (int rc = launch_transport(h, n, nullptr, ion, seed, first_index, records != nullptr))
    return rc;
  if (int rc = sync_and_check(h))
    return rc;
  if (records && n)
    return mtb_fetch_records(h, n, records);
  return MTB_OK;
}

int
mtb_reset_tallies(mtb_handle * h)
{
  if (int rc = ensure_ready(h))
    return rc;
  if (int rc = drain(h))
    return rc;
  MTB_CUDA(cudaMemsetAsync(h->d_u64.p, 0, h->u64_size * sizeof(unsigned long long), h->stream));
  MTB_CUDA(cudaMemsetAsync(h->d_f64.p, 0, 2 * sizeof(double), h->stream));
  MTB_CUDA(cudaStreamSynchronize(h->stream));
  return MTB_OK;
}

int
mtb_get_counters(mtb_handle * h, mtb_counters * out)
{
  if (int rc = ensure_ready(h))
    return rc;
  if (!out)
    return fail(MTB_EINVAL, "null argument");
  unsigned long long c[CNT_COUNT];
  double f[2];
  if (int rc = drain(h))
    return rc;
  MTB_CUDA(cudaMemcpy(c, h->d_u64.p, sizeof(c), cudaMemcpyDeviceToHost));
  MTB_CUDA(cudaMemcpy(f, h->d_f64.p, sizeof(f), cudaMemcpyDeviceToHost));
  out->vacancies_created = c[CNT_VAC];
  out->replacements = c[CNT_REPL];
  out->steps = c[CNT_STEPS];
  out->ions = c[CNT_IONS];
  out->primaries = c[CNT_PRIMARIES];
  out->recoils_queued = c[CNT_QUEUED];
  out->lost = c[CNT_LOST];
  out->left_sample = c[CNT_LEFT];
  out->hist_clamped = c[CNT_CLAMPED];
  out->stack_max = c[CNT_STACKMAX];
  out->EelTotal = f[0];
  out->EnucTotal = f[1];
  return MTB_OK;
}

int
mtb_hist_bins(mtb_handle * h, size_t * bins, size_t * evac_rows)
{
  if (int rc = ensure_ready(h))
    return rc;
  if (bins)
    *bins = (size_t)h->P.hist_bins;
  if (evac_rows)
    *evac_rows = (size_t)h->P.evac_rows;
  return MTB_OK;
}

int
mtb_get_vac_depth(mtb_handle * h, uint64_t * vac, uint64_t * repl, size_t capacity, size_t * n_bins)
{
  if (int rc = ensure_ready(h))
    return rc;
  const size_t B = (size_t)h->P.hist_bins;
  std::vector<unsigned long long> buf(2 * B);
  if (int rc = drain(h))
    return rc;
  MTB_CUDA(cudaMemcpy(buf.data(), h->d_u64.p + off_vac(h->P), 2 * B * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  size_t last = 0;
  for (size_t i = 0; i < B; ++i)
    if (buf[i] || buf[B + i])
      last = i + 1;
  if (n_bins)
    *n_bins = last;
  for (size_t i = 0; i < capacity; ++i)
  {
    if (vac)
      vac[i] = i < B ? buf[i] : 0;
    if (repl)
      repl[i] = i < B ? buf[B + i] : 0;
  }
  return last > capacity ? fail(MTB_ECAPACITY, "histogram longer than the output buffer") : MTB_OK;
}

int
mtb_get_vac_energy(mtb_handle * h, uint64_t * evac, size_t rows, size_t bins)
{
  if (int rc = ensure_ready(h))
    return rc;
  if (!(h->host.cfg.tally_mask & MTB_TALLY_VAC_ENERGY) || !evac)
    return fail(MTB_EINVAL, "MTB_TALLY_VAC_ENERGY not enabled");
  const size_t B = (size_t)h->P.hist_bins, R = (size_t)h->P.evac_rows;
  std::vector<unsigned long long> buf(R * B);
  if (int rc = drain(h))
    return rc;
  MTB_CUDA(cudaMemcpy(buf.data(), h->d_u64.p + off_evac(h->P), R * B * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  std::memset(evac, 0, rows * bins * sizeof(uint64_t));
  for (size_t r = 0; r < std::min(R, rows); ++r)
    for (size_t x = 0; x < std::min(B, bins); ++x)
      evac[r * bins + x] = buf[r * B + x];
  return MTB_OK;
}

int
mtb_get_vacmap(mtb_handle * h, uint64_t * vmap)
{
  if (int rc = ensure_ready(h))
    return rc;
  if (!vmap)
    return fail(MTB_EINVAL, "null argument");
  if (int rc = drain(h))
    return rc;
  MTB_CUDA(cudaMemcpy(vmap, h->d_u64.p + off_vmap(h->P), MTB_VMAP_NX * MTB_VMAP_NY * 3 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return MTB_OK;
}

int
mtb_get_range_list(mtb_handle * h, float * x, int32_t * Z, size_t capacity, size_t * n)
{
  if (int rc = ensure_ready(h))
    return rc;
  if (!(h->host.cfg.tally_mask & MTB_TALLY_RANGE))
    return fail(MTB_EINVAL, "MTB_TALLY_RANGE not enabled");
  unsigned long long cnt = 0;
  if (int rc = drain(h))
    return rc;
  MTB_CUDA(cudaMemcpy(&cnt, h->d_u64.p + CNT_RANGE_N, sizeof(cnt), cudaMemcpyDeviceToHost));
  if (n)
    *n = (size_t)cnt;
  const size_t have = (size_t)std::min<unsigned long long>(cnt, h->P.range_cap);
  const size_t take = std::min(have, capacity);
  std::vector<RangeEntry> buf(take);
  if (take)
    MTB_CUDA(cudaMemcpy(buf.data(), h->d_range.p, take * sizeof(RangeEntry), cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < take; ++i)
  {
    if (x)
      x[i] = buf[i].x;
    if (Z)
      Z[i] = buf[i].Z;
  }
  if (cnt > h->P.range_cap)
    return fail(MTB_ECAPACITY, "range list overflowed range_capacity");
  return cnt > capacity ? fail(MTB_ECAPACITY, "range list longer than the output buffer") : MTB_OK;
}

int
mtb_get_ion_log(mtb_handle * h, mtb_ion_log * out, size_t capacity, size_t * n)
{
  if (int rc = ensure_ready(h))
    return rc;
  if (!(h->host.cfg.tally_mask & MTB_TALLY_IONLOG))
    return fail(MTB_EINVAL, "MTB_TALLY_IONLOG not enabled");
  unsigned long long cnt = 0;
  if (int rc = drain(h))
    return rc;
  MTB_CUDA(cudaMemcpy(&cnt, h->d_u64.p + CNT_IONLOG_N, sizeof(cnt), cudaMemcpyDeviceToHost));
  const size_t have = (size_t)std::min<unsigned long long>(cnt, h->P.ionlog_cap);
  std::vector<mtb_ion_log> raw(have);
  if (have)
    MTB_CUDA(cudaMemcpy(raw.data(), h->d_ionlog.p, have * sizeof(mtb_ion_log), cudaMemcpyDeviceToHost));
  // join birth halves (state == -1) with death halves by stream id
  std::unordered_map<uint64_t, size_t> birth;
  birth.reserve(have);
  for (size_t i = 0; i < have; ++i)
    if (raw[i].state == -1)
      birth[raw[i].uid] = i;
  size_t m = 0;
  for (size_t i = 0; i < have; ++i)
  {
    if (raw[i].state == -1)
      continue;
    auto it = birth.find(raw[i].uid);
    if (it == birth.end())
      continue;
    if (m < capacity && out)
    {
      mtb_ion_log e = raw[i];
      std::memcpy(e.pos0, raw[it->second].pos0, sizeof(e.pos0));
      e.E0 = raw[it->second].E0;
      out[m] = e;
    }
    ++m;
  }
  if (n)
    *n = m;
  if (cnt > h->P.ionlog_cap)
    return fail(MTB_ECAPACITY, "ion log overflowed ionlog_capacity");
  return m > capacity ? fail(MTB_ECAPACITY, "ion log longer than the output buffer") : MTB_OK;
}

int
mtb_clear_lists(mtb_handle * h)
{
  if (int rc = ensure_ready(h))
    return rc;
  if (int rc = drain(h))
    return rc;
  MTB_CUDA(cudaMemsetAsync(h->d_u64.p + CNT_IONLOG_N, 0, sizeof(unsigned long long), h->stream));
  MTB_CUDA(cudaMemsetAsync(h->d_u64.p + CNT_RANGE_N, 0, sizeof(unsigned long long), h->stream));
  MTB_CUDA(cudaStreamSynchronize(h->stream));
  return MTB_OK;
}

int
mtb_tally_device_views(mtb_handle * h, void ** u64_dev, size_t * n_u64, void ** f64_dev, size_t * n_f64)
{
  if (int rc = ensure_ready(h))
    return rc;
  if (int rc = drain(h))
    return rc;
  if (u64_dev)
    *u64_dev = h->d_u64.p;
  if (n_u64)
    *n_u64 = h->u64_size;
  if (f64_dev)
    *f64_dev = h->d_f64.p;
  if (n_f64)
    *n_f64 = 2;
  return MTB_OK;
}

int
mtb_trim_one(mtb_handle * h, mtb_ion * ion, uint64_t seed, uint64_t uid, int32_t * final_state,
             mtb_event * events, size_t capacity, size_t * n_events)
{
  if (!h || !ion)
    return fail(MTB_EINVAL, "null argument");
  if (h->have_materials && register_primary_species(h->host, 1, ion))
    h->dirty = true;
  if (int rc = ensure_ready(h))
    return rc;
  const size_t cap = std::max<size_t>(capacity, 1);
  MTB_CUDA(h->d_events.ensure(cap));
  LaunchParams P = h->P;
  P.primaries = nullptr;
  P.beam = *ion;
  P.n_primaries = 1;
  P.first_index = 0;
  P.single_uid = uid;
  P.key0 = (uint32_t)seed;
  P.key1 = (uint32_t)(seed >> 32);
  philox_round_keys(P.key0, P.key1, P.rk);
  P.share_min_E = h->share_min_E;
  P.records = nullptr;
  P.events = h->d_events.p;
  P.events_cap = events ? capacity : 0;
  MTB_CUDA(h->d_custom_rows.ensure((size_t)(2 + P.n_materials + P.n_tclass)));
  P.custom_rows = h->d_custom_rows.p;
  P.tally_mask = 0; // hooks run on the host in this mode
  MTB_CUDA(cudaMemsetAsync(&P.u64[CNT_EVENTS_N], 0, sizeof(unsigned long long), h->stream));
  trim_one_kernel<<<1, 32, h->smem_bytes, h->stream>>>(P);
  MTB_CUDA(cudaGetLastError());
  unsigned long long cnt = 0;
  MTB_CUDA(cudaMemcpyAsync(&cnt, h->d_u64.p + CNT_EVENTS_N, sizeof(cnt), cudaMemcpyDeviceToHost, h->stream));
  MTB_CUDA(cudaStreamSynchronize(h->stream));
  if (int rc = check_skipped_ions(h))
    return rc;
  const size_t take = (size_t)std::min<unsigned long long>(cnt, P.events_cap);
  mtb_event last;
  bool have_last = false;
  if (take)
  {
    MTB_CUDA(cudaMemcpy(events, h->d_events.p, take * sizeof(mtb_event), cudaMemcpyDeviceToHost));
    if (take == cnt)
    {
      last = events[take - 1];
      have_last = true;
    }
  }
  if (n_events)
    *n_events = (size_t)cnt;
  if (cnt > P.events_cap)
    return fail(MTB_ECAPACITY, "event buffer too small");
  if (have_last)
  {
    std::memcpy(ion->pos, last.pka_pos, sizeof(ion->pos));
    std::memcpy(ion->dir, last.pka_dir, sizeof(ion->dir));
    ion->E = last.pka_E;
    if (final_state)
      *final_state = last.pka_state;
  }
  else if (final_state)
    *final_state = MTB_MOVING;
  return MTB_OK;
}

int
mtb_trim_many(mtb_handle * h, size_t n, mtb_ion * ions, uint64_t seed, uint64_t first_uid, const uint64_t * uids,
              int32_t * final_states, mtb_event * events, size_t events_per_ion, uint32_t * counts)
{
  if (!h || (n && (!ions || !events || !counts)) || events_per_ion < 1)
    return fail(MTB_EINVAL, "bad argument");
  if (!n)
    return MTB_OK;
  if (h->have_materials && register_primary_species(h->host, std::min<size_t>(n, 64), ions))
    h->dirty = true;
  if (int rc = ensure_ready(h))
    return rc;
  if (int rc = drain(h))
    return rc;
  const size_t lanes = (n + 31) / 32 * 32;
  MTB_CUDA(h->d_events.ensure(n * events_per_ion));
  MTB_CUDA(h->d_event_counts.ensure(n));
  // the event block comes back in one copy, slots past an ion's last event included: keep them defined
  MTB_CUDA(cudaMemsetAsync(h->d_events.p, 0, n * events_per_ion * sizeof(mtb_event), h->stream));
  MTB_CUDA(h->d_primaries.upload(ions, n, h->stream));
  h->n_resident = 0; // the resident primaries of mtb_upload_primaries were replaced
  LaunchParams P = h->P;
  P.primaries = h->d_primaries.p;
  P.n_primaries = n;
  P.first_index = 0;
  P.single_uid = first_uid;
  P.uid_list = nullptr;
  if (uids)
  {
    MTB_CUDA(h->d_uids.upload(uids, n, h->stream));
    P.uid_list = h->d_uids.p;
  }
  P.key0 = (uint32_t)seed;
  P.key1 = (uint32_t)(seed >> 32);
  philox_round_keys(P.key0, P.key1, P.rk);
  P.share_min_E = h->share_min_E;
  P.records = nullptr;
  P.index_list = nullptr;
  P.deferred = nullptr;
  P.events = h->d_events.p;
  P.events_cap = events_per_ion;
  P.event_counts = h->d_event_counts.p;
  MTB_CUDA(h->d_custom_rows.ensure(lanes * (size_t)(2 + P.n_materials + P.n_tclass)));
  P.custom_rows = h->d_custom_rows.p;
  P.tally_mask = 0; // hooks run on the host in this mode
  trim_one_kernel<<<(unsigned)(lanes / 32), 32, h->smem_bytes, h->stream>>>(P);
  MTB_CUDA(cudaGetLastError());
  MTB_CUDA(cudaMemcpyAsync(counts, h->d_event_counts.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  MTB_CUDA(cudaStreamSynchronize(h->stream));
  if (int rc = check_skipped_ions(h))
    return rc;
  // one copy of the used part of the event block: up to the last ion that has any event
  size_t last = 0;
  for (size_t i = 0; i < n; ++i)
    if (counts[i])
      last = i + 1;
  if (last)
    MTB_CUDA(cudaMemcpy(events, h->d_events.p, last * events_per_ion * sizeof(mtb_event), cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < n; ++i)
  {
    int32_t st = MTB_MOVING; // no collision: the ion started in vacuum (trim.C:80-82)
    if (counts[i] && counts[i] <= events_per_ion)
    {
      const mtb_event & e = events[i * events_per_ion + counts[i] - 1];
      std::memcpy(ions[i].pos, e.pka_pos, sizeof(ions[i].pos));
      std::memcpy(ions[i].dir, e.pka_dir, sizeof(ions[i].dir));
      ions[i].E = e.pka_E;
      st = e.pka_state;
    }
    if (final_states)
      final_states[i] = st;
  }
  return MTB_OK;
}

int
mtb_stopping(mtb_handle * h, int material, size_t n, const int32_t * Z1, const double * m1, const double * E,
             double * out)
{
  if (int rc = ensure_ready(h))
    return rc;
  if (material < 0 || material >= (int)h->mat_map.size() || !Z1 || !m1 || !E || !out)
    return fail(MTB_EINVAL, "bad argument");
  material = h->mat_map[material]; // identical materials of a layer stack are folded into one on the device
  for (size_t i = 0; i < n; ++i)
    if (Z1[i] < 1 || Z1[i] > MTB_NZ)
      return fail(MTB_EINVAL, "projectile Z out of range");
  if (!n)
    return MTB_OK;
  DevBuf<int32_t> dZ;
  DevBuf<double> dm, dE, dout;
  MTB_CUDA(dZ.upload(Z1, n, h->stream));
  MTB_CUDA(dm.upload(m1, n, h->stream));
  MTB_CUDA(dE.upload(E, n, h->stream));
  MTB_CUDA(dout.ensure(n));
  const unsigned blocks = (unsigned)((n + 127) / 128);
  stopping_kernel<<<blocks, 128, h->smem_bytes, h->stream>>>(h->P, material, n, dZ.p, dm.p, dE.p, dout.p);
  MTB_CUDA(cudaGetLastError());
  MTB_CUDA(cudaMemcpyAsync(out, dout.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  MTB_CUDA(cudaStreamSynchronize(h->stream));
  return MTB_OK;
}

int
mtb_measure_fp32_peak(int device, double * tflops, float * ms_out)
{
  if (!tflops)
    return fail(MTB_EINVAL, "null argument");
  MTB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  MTB_CUDA(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
  float * d = nullptr;
  MTB_CUDA(cudaMalloc(&d, (size_t)blocks * threads * sizeof(float)));
  cudaEvent_t e0, e1;
  MTB_CUDA(cudaEventCreate(&e0));
  MTB_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep)
  {
    MTB_CUDA(cudaEventRecord(e0));
    fp32_peak_kernel<<<blocks, threads>>>(d, iters, 0.999f, 0.001f);
    MTB_CUDA(cudaEventRecord(e1));
    MTB_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    MTB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep >= 1 && ms < best)
      best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  const double flops = (double)blocks * threads * (double)iters * 16.0 * 8.0 * 2.0;
  *tflops = flops / (best * 1e-3) / 1e12;
  if (ms_out)
    *ms_out = best;
  return MTB_OK;
}

// Single-process multi-GPU tally reduction (one handle per GPU): ncclAllReduce(sum) over the additive
// u64 block and the f64 block, ncclAllReduce(max) over the stack high-water mark.  NCCL is loaded
// with dlopen so the library has no link-time dependency on it; multi-process jobs (one rank per
// GPU) reduce the same device blocks through torch.distributed instead (mytrim_b200/dist.py).
namespace
{
struct NcclApi
{
  void * lib = nullptr;
  int (*CommInitAll)(void **, int, const int *) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char * (*GetErrorString)(int) = nullptr;
  bool load()
  {
    if (lib)
      return true;
    // An NCCL the process has already loaded wins (a later `import torch` must not find an older libnccl.so.2 under
    // the same soname than the one it was built against); then $MYTRIM_B200_NCCL_LIB (capi.py points it at the NCCL
    // bundled with torch); then the system library.
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!lib)
      if (const char * env = std::getenv("MYTRIM_B200_NCCL_LIB"))
        lib = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!lib)
      for (const char * name : {"libnccl.so.2", "libnccl.so"})
        if ((lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL)))
          break;
    if (!lib)
      return false;
    CommInitAll = (int (*)(void **, int, const int *))dlsym(lib, "ncclCommInitAll");
    CommDestroy = (int (*)(void *))dlsym(lib, "ncclCommDestroy");
    AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(lib, "ncclAllReduce");
    GroupStart = (int (*)())dlsym(lib, "ncclGroupStart");
    GroupEnd = (int (*)())dlsym(lib, "ncclGroupEnd");
    GetErrorString = (const char * (*)(int))dlsym(lib, "ncclGetErrorString");
    return CommInitAll && CommDestroy && AllReduce && GroupStart && GroupEnd;
  }
};
NcclApi g_nccl;
constexpr int kNcclInt64 = 4, kNcclUint64 = 5, kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2;
} // namespace

int
mtb_allreduce(mtb_handle ** handles, int n_handles)
{
  if (!handles || n_handles < 1)
    return fail(MTB_EINVAL, "no handles");
  if (n_handles == 1)
    return MTB_OK;
  for (int i = 0; i < n_handles; ++i)
  {
    if (int rc = ensure_ready(handles[i]))
      return rc;
    if (handles[i]->u64_size != handles[0]->u64_size)
      return fail(MTB_EINVAL, "handles have different tally layouts");
    if (int rc = drain(handles[i]))
      return rc;
  }
  if (!g_nccl.load())
    return fail(MTB_ENCCL, "libnccl.so.2 not found");
  std::vector<void *> comms(n_handles, nullptr);
  std::vector<int> devs(n_handles);
  for (int i = 0; i < n_handles; ++i)
    devs[i] = handles[i]->device;
  int rc = g_nccl.CommInitAll(comms.data(), n_handles, devs.data());
  if (rc != 0)
    return fail(MTB_ENCCL, std::string("ncclCommInitAll: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
  const size_t nhist = handles[0]->u64_size - CNT_COUNT;
  rc = g_nccl.GroupStart();
  for (int i = 0; i < n_handles && rc == 0; ++i)
  {
    mtb_handle * h = handles[i];
    cudaSetDevice(h->device);
    rc = g_nccl.AllReduce(h->d_u64.p, h->d_u64.p, CNT_STACKMAX, kNcclUint64, kNcclSum, comms[i], h->stream);
    if (rc == 0)
      rc = g_nccl.AllReduce(h->d_u64.p + CNT_STACKMAX, h->d_u64.p + CNT_STACKMAX, 1, kNcclUint64, kNcclMax, comms[i], h->stream);
    if (rc == 0)
      rc = g_nccl.AllReduce(h->d_u64.p + CNT_COUNT, h->d_u64.p + CNT_COUNT, nhist, kNcclUint64, kNcclSum, comms[i], h->stream);
    if (rc == 0)
      rc = g_nccl.AllReduce(h->d_f64.p, h->d_f64.p, 2, kNcclFloat64, kNcclSum, comms[i], h->stream);
  }
  const int rc_end = g_nccl.GroupEnd();
  if (rc == 0)
    rc = rc_end;
  for (int i = 0; i < n_handles; ++i)
  {
    cudaSetDevice(handles[i]->device);
    cudaStreamSynchronize(handles[i]->stream);
    g_nccl.CommDestroy(comms[i]);
  }
  if (rc != 0)
    return fail(MTB_ENCCL, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
  (void)kNcclInt64;
  return MTB_OK;
}

} // extern "C"
