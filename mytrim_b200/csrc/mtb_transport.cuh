// mtb_transport.cuh — the cascade loop of one lane (one CUDA thread).
//
// Replaces, for the primaries a lane draws from the global work counter, the reference's
//   queue.push(pka); while(!queue.empty()){ r=pop; sample->averages(r); trim->trim(r,queue); }
// (apps/runmytrim.C:76-92) together with the body of TrimBase::trim (trim.C:35-425).
//
// B200-first structure: a lane owns a whole cascade.  The ion in flight lives in registers
// (position and energy in FP64 — cheap on B200 — everything else FP32); suspended ions live on a
// lane-private 64-byte-per-entry stack in HBM/L2.  When a collision leaves two moving ions the
// lane keeps the one with LESS energy and pushes the other, which bounds the stack depth by
// log2(E0/E_threshold) <= 32 and never touches memory when the projectile stops in the same
// collision.  Every ion draws its randoms from its own Philox4x32-10 stream keyed by a
// scheduling-independent id, so the traversal order (depth first here, FIFO in the reference and
// in the oracle) does not change any result.
//
// The file compiles for the host too (MTB_HOSTSIM): tests/hostsim.cpp drives the same loop
// single-threaded to debug control flow without a GPU.  It is NOT a product fallback.
#ifndef MTB_TRANSPORT_CUH
#define MTB_TRANSPORT_CUH

#include "mtb_physics.cuh"

namespace mtb
{

#if MTB_DEVICE_CODE
#define MTB_ATOMIC_ADD(ptr, val) atomicAdd((ptr), (val))
#define MTB_ATOMIC_MAX(ptr, val) atomicMax((ptr), (val))
// warp-uniform loop exit: all 32 lanes vote every iteration, which is also where the warp reconverges
#define MTB_WARP_ALL(pred) __all_sync(0xffffffffu, (pred))
#define MTB_WARP_ANY(pred) __any_sync(0xffffffffu, (pred))
#else
#define MTB_WARP_ALL(pred) (pred)
#define MTB_WARP_ANY(pred) (pred)
template <class T, class U>
inline T
host_fetch_add(T * p, U v)
{
  T old = *p;
  *p += (T)v;
  return old;
}
template <class T, class U>
inline T
host_fetch_max(T * p, U v)
{
  T old = *p;
  if ((T)v > old)
    *p = (T)v;
  return old;
}
#define MTB_ATOMIC_ADD(ptr, val) ::mtb::host_fetch_add((ptr), (val))
#define MTB_ATOMIC_MAX(ptr, val) ::mtb::host_fetch_max((ptr), (val))
#endif

// Add to a 64-bit block accumulator in shared memory.  Shared memory has no native 64-bit add (the
// compiler emits a compare-and-swap loop of a dozen instructions per site); two native 32-bit adds
// with a carry do the same.  The accumulators are only read after the CTA has synchronised.
MTB_HD void
block_add(unsigned long long * acc, uint32_t v)
{
#if MTB_DEVICE_CODE
  unsigned int * w = reinterpret_cast<unsigned int *>(acc);
  const unsigned int old = atomicAdd(w, v);
  if (old + v < old)
    atomicAdd(w + 1, 1u);
#else
  *acc += v;
#endif
}

// Compile-time specialisation of the lane loop.  A variant is a set of features (which run-time
// options it reads from LaunchParams) plus the tallies it can produce; everything outside the set is
// not even compiled in.  That matters twice: fewer instructions per collision, and a collision loop
// that fits the instruction cache (the all-features loop spills out of the 32 KB L1.5 on the uo2
// workload: "no instruction" was its top stall reason).  pick_variant() chooses the leanest variant
// that covers a configuration:
//   FAST      north-star options: UNIVERSAL potential, follow ALL, vacancies_created++, TrimVacCount depth
//             tallies, solid/layered sample, no CUT boundaries, projectile classes only
//   MONO      FAST for a sample of one single-element material (Cu->Cu, H->Fe, He->Fe, C->W of BASELINE.json)
//   CLUSTERS  the tests/uo2 shape: sampleClusters geometry, per-primary species (fission fragments),
//             ion log / energy partition, otherwise the north-star options
//   LAYERS    solid/layered sample with any follow policy, vacancy model and tally (vacenergycount, range,
//             TrimPrimaries/Recoils, PhononOut, VacMap): the other runmytrim / mytrim_layers shapes
//   GENERIC   every option at run time
enum Feature : uint32_t
{
  F_EVENTS = 1u << 0,    // single-ion mode behind mtb_trim_one: every collision is reported, no recoil is followed
  F_SHARE = 1u << 1,     // lanes without work adopt suspended ions of other lanes of the CTA
  F_CUSTOM = 1u << 2,    // primaries whose (Z, m) has no projectile class build private table rows
  F_CLUSTERS = 1u << 3,  // sampleClusters geometry
  F_GEOM_ANY = 1u << 4,  // SampleWire, SampleBurriedWire
  F_CUT = 1u << 5,       // CUT boundary conditions (trim.C:344-352)
  F_POTENTIAL = 1u << 6, // potential chosen at run time (else UNIVERSAL)
  F_FOLLOW = 1u << 7,    // follow policy chosen at run time (else ALL)
  F_VACMODEL = 1u << 8,  // vacancy model chosen at run time (else vacancies_created++)
  F_TALLY_RT = 1u << 9,  // tallies switched at run time within kTally (else exactly kTally)
  F_DIAG = 1u << 10,     // stack high-water mark
  F_MONO = 1u << 11,     // the sample is one material made of one element (solid or layers of it): no geometry
                         // look-up, no vacuum test, no target pick, no loop over elements in the stopping
  F_NOREC = 1u << 12     // the launch asked for no per-primary records: the record flag of the ion word, its test on
                         // every exit of an ion and the record writes are compiled out (3 registers less: 8 CTAs/SM)
};

template <uint32_t F, uint32_t TALLY>
struct TraitsT
{
  static constexpr uint32_t kF = F, kTally = TALLY;
  static constexpr bool kEvents = (F & F_EVENTS) != 0, kShare = (F & F_SHARE) != 0, kCustom = (F & F_CUSTOM) != 0;
  MTB_HD static constexpr bool has(uint32_t f) { return (F & f) != 0; }
};

constexpr uint32_t kFeatFast = 0;
constexpr uint32_t kTallyFast = MTB_TALLY_VAC_DEPTH;
constexpr uint32_t kFeatMono = F_MONO;
constexpr uint32_t kFeatClusters = F_CUSTOM | F_CLUSTERS | F_TALLY_RT;
constexpr uint32_t kTallyClusters = MTB_TALLY_IONLOG | MTB_TALLY_PHONON;
// the tests/uo2 driver itself asks for the ion log and nothing else: tallies fixed at compile time (no run-time
// tally tests on the hand-over paths, no energy-partition sums), measured -3.6 % on the uo2 workload (r02v)
constexpr uint32_t kFeatClustersLog = F_CUSTOM | F_CLUSTERS;
constexpr uint32_t kTallyClustersLog = MTB_TALLY_IONLOG;
constexpr uint32_t kFeatLayers = F_FOLLOW | F_VACMODEL | F_TALLY_RT;
constexpr uint32_t kFeatGeneric =
    F_CUSTOM | F_CLUSTERS | F_GEOM_ANY | F_CUT | F_POTENTIAL | F_FOLLOW | F_VACMODEL | F_TALLY_RT | F_DIAG;
constexpr uint32_t kTallyAll = 0xffffffffu;

typedef TraitsT<kFeatFast, kTallyFast> TraitsFast;
typedef TraitsT<kFeatFast | F_SHARE, kTallyFast> TraitsFastShare;
typedef TraitsT<kFeatMono, kTallyFast> TraitsMono;
typedef TraitsT<kFeatMono | F_SHARE, kTallyFast> TraitsMonoShare;
typedef TraitsT<kFeatMono | F_NOREC, kTallyFast> TraitsMonoNoRec; // what bench.py times: TrimVacCount tallies only
// MONO with the TrimVacEnergyCount tally instead (validation/c_on_w/input.json: "type": "vacenergycount")
typedef TraitsT<kFeatMono, MTB_TALLY_VAC_ENERGY> TraitsMonoEvac;
typedef TraitsT<kFeatMono | F_SHARE, MTB_TALLY_VAC_ENERGY> TraitsMonoEvacShare;
// FAST with the energy partition of TrimPhononOut instead of the depth histograms
typedef TraitsT<kFeatFast, MTB_TALLY_PHONON> TraitsFastPhonon;
typedef TraitsT<kFeatFast | F_SHARE, MTB_TALLY_PHONON> TraitsFastPhononShare;
typedef TraitsT<kFeatClusters, kTallyClusters> TraitsClusters;
typedef TraitsT<kFeatClusters | F_SHARE, kTallyClusters> TraitsClustersShare;
typedef TraitsT<kFeatClustersLog, kTallyClustersLog> TraitsClustersLog;
typedef TraitsT<kFeatClustersLog | F_SHARE, kTallyClustersLog> TraitsClustersLogShare;
// LAYERS without any tally beyond counters and records: what TrimPrimaries / TrimRecoils (apps/mytrim_layers.C) ask for
constexpr uint32_t kFeatLayersPlain = F_FOLLOW | F_VACMODEL;
typedef TraitsT<kFeatLayersPlain, 0u> TraitsLayersPlain;
typedef TraitsT<kFeatLayersPlain | F_SHARE, 0u> TraitsLayersPlainShare;
typedef TraitsT<kFeatLayers, kTallyAll> TraitsLayers;
typedef TraitsT<kFeatLayers | F_SHARE, kTallyAll> TraitsLayersShare;
typedef TraitsT<kFeatGeneric, kTallyAll> TraitsGeneric;
typedef TraitsT<kFeatGeneric | F_SHARE, kTallyAll> TraitsGenericShare;
typedef TraitsT<kFeatGeneric | F_EVENTS, kTallyAll> TraitsEvents;

template <class TR>
MTB_HD bool
tally_on(const LaunchParams & P, uint32_t bit)
{
  return (TR::kTally & bit) != 0 && (!TR::has(F_TALLY_RT) || (P.tally_mask & bit) != 0);
}

enum Variant
{
  VARIANT_FAST = 0,
  VARIANT_CLUSTERS,
  VARIANT_LAYERS,
  VARIANT_GENERIC,
  VARIANT_MONO,
  VARIANT_MONO_NOREC, // chosen per launch (mtb_engine.cu: launch_transport), never by pick_variant
  VARIANT_CLUSTERS_LOG,
  VARIANT_MONO_EVAC,
  VARIANT_FAST_PHONON,
  VARIANT_LAYERS_PLAIN,
  VARIANT_COUNT
};

inline const char *
variant_name(Variant v)
{
  switch (v)
  {
    case VARIANT_FAST: return "FAST";
    case VARIANT_CLUSTERS: return "CLUSTERS";
    case VARIANT_LAYERS: return "LAYERS";
    case VARIANT_GENERIC: return "GENERIC";
    case VARIANT_MONO: return "MONO";
    case VARIANT_MONO_NOREC: return "MONO-NOREC";
    case VARIANT_CLUSTERS_LOG: return "CLUSTERS-LOG";
    case VARIANT_MONO_EVAC: return "MONO-EVAC";
    case VARIANT_FAST_PHONON: return "FAST-PHONON";
    case VARIANT_LAYERS_PLAIN: return "LAYERS-PLAIN";
    default: return "?";
  }
}

inline uint32_t
variant_features(Variant v)
{
  return (v == VARIANT_FAST || v == VARIANT_FAST_PHONON) ? kFeatFast : (v == VARIANT_MONO || v == VARIANT_MONO_EVAC) ? kFeatMono : v == VARIANT_MONO_NOREC ? (kFeatMono | F_NOREC) : v == VARIANT_CLUSTERS ? kFeatClusters : v == VARIANT_CLUSTERS_LOG ? kFeatClustersLog : v == VARIANT_LAYERS ? kFeatLayers : v == VARIANT_LAYERS_PLAIN ? kFeatLayersPlain : kFeatGeneric;
}

// Features a configuration needs (F_CUSTOM is decided per primary: variants without it hand
// class-less primaries to a second launch of a variant that has it).
inline uint32_t
needed_features(const LaunchParams & P)
{
  uint32_t f = 0;
  if (P.geom_kind == MTB_GEOM_CLUSTERS)
    f |= F_CLUSTERS;
  else if (P.geom_kind != MTB_GEOM_SOLID && P.geom_kind != MTB_GEOM_LAYERS)
    f |= F_GEOM_ANY;
  if (P.bc[0] == MTB_BC_CUT || P.bc[1] == MTB_BC_CUT || P.bc[2] == MTB_BC_CUT)
    f |= F_CUT;
  // outside a non-periodic clusters box the lookup returns vacuum: same code path as CUT-less generic
  if (P.potential != MTB_POT_UNIVERSAL)
    f |= F_POTENTIAL;
  if (P.follow != MTB_FOLLOW_ALL)
    f |= F_FOLLOW;
  if (P.vacancy_model != MTB_VAC_COUNT)
    f |= F_VACMODEL;
  return f;
}

inline bool
variant_covers(uint32_t feat, uint32_t tally, const LaunchParams & P)
{
  const uint32_t want = P.tally_mask & ~(uint32_t)MTB_TALLY_RECORDS;
  if (needed_features(P) & ~feat)
    return false;
  return (feat & F_TALLY_RT) ? (want & ~tally) == 0 : want == tally;
}

// The leanest variant that covers the configuration; `custom` = the launch may contain primaries
// without a projectile class and cannot defer them (second launch of a deferral, beam mode).
inline Variant
pick_variant(const LaunchParams & P, bool custom)
{
  if (!custom && variant_covers(kFeatFast, kTallyFast, P))
    return P.mono ? VARIANT_MONO : VARIANT_FAST;
  if (!custom && P.mono && variant_covers(kFeatMono & ~(uint32_t)F_MONO, MTB_TALLY_VAC_ENERGY, P))
    return VARIANT_MONO_EVAC;
  if (!custom && variant_covers(kFeatFast, MTB_TALLY_PHONON, P))
    return VARIANT_FAST_PHONON;
  if (P.geom_kind == MTB_GEOM_CLUSTERS && variant_covers(kFeatClustersLog, kTallyClustersLog, P))
    return VARIANT_CLUSTERS_LOG;
  if (variant_covers(kFeatClusters, kTallyClusters, P))
    return VARIANT_CLUSTERS;
  if (!custom && variant_covers(kFeatLayersPlain, 0u, P))
    return VARIANT_LAYERS_PLAIN;
  if (!custom && variant_covers(kFeatLayers, kTallyAll, P))
    return VARIANT_LAYERS;
  return VARIANT_GENERIC;
}

// Does this configuration qualify for TraitsFast?  (The launcher additionally requires that every
// primary species has a projectile class; per-primary masses go through a variant with F_CUSTOM.)
inline bool
fast_path_ok(const LaunchParams & P)
{
  return variant_covers(kFeatFast, kTallyFast, P);
}

// Block-local views: the small tables staged in shared memory plus block accumulators.
struct BlockCtx
{
  const DevElement * elements;
  const DevMaterial * materials;
  const DevIonZ * ionz;
  const LowStop * lowstop;
  const ProjClass * pclass;
  const PairM * pairm;
  const PairE * paire;
  const double * layer_cum;
  const int32_t * layer_mat;
  unsigned int * hist_vac;  // [smem_hist_bins] or null
  unsigned int * hist_repl; // [smem_hist_bins] or null; directly behind hist_vac
  unsigned long long * blk_u64; // [CNT_COUNT]
  double * blk_f64;             // [2]
  PoolSlot * pool;               // [MTB_POOL_SLOTS] work-sharing ring of this CTA (share kernels)
  unsigned long long * pool_ctl; // [POOL_CTL_COUNT]
};

MTB_HD size_t
off_vac(const LaunchParams &)
{
  return CNT_COUNT;
}
MTB_HD size_t
off_repl(const LaunchParams & P)
{
  return CNT_COUNT + (size_t)P.hist_bins;
}
MTB_HD size_t
off_evac(const LaunchParams & P)
{
  return CNT_COUNT + 2 * (size_t)P.hist_bins;
}
MTB_HD size_t
off_vmap(const LaunchParams & P)
{
  return off_evac(P) + ((P.tally_mask & MTB_TALLY_VAC_ENERGY) ? (size_t)P.evac_rows * (size_t)P.hist_bins : 0);
}
MTB_HD size_t
u64_block_size(const LaunchParams & P)
{
  return off_vmap(P) + MTB_VMAP_NX * MTB_VMAP_NY * 3;
}

// ---------------------------------------------------------------------------------------------
// geometry: the lookupMaterial() family.  Returns the de-duplicated
// material id, or -1 for vacuum; *cluster receives the cluster index (clusters geometry).
// ---------------------------------------------------------------------------------------------
// j mod kn for a cell index at most one period outside [0, kn) (the scan range is [k - ks, k + ks]
// with 0 <= k < kn); the general modulo only when the neighbourhood is wider than the hash itself
MTB_HD int
wrap_cell(int j, int kn)
{
  if (j < 0)
    j += kn;
  else if (j >= kn)
    j -= kn;
  if (j < 0 || j >= kn)
  {
    j %= kn;
    if (j < 0)
      j += kn;
  }
  return j;
}

MTB_HD_COLD int
lookup_cluster(const LaunchParams & P, double px, double py, double pz, float * safe)
{
  *safe = 0.0f;
  // sampleClusters::lookupCluster(pos, 0) — sample_clusters.C:59-133
  const double pos[3] = {px, py, pz};
  int kc[3], k1[3], k2[3];
  for (int i = 0; i < 3; ++i)
  {
    // cell index floor((pos * kn) / w): a multiplication by kn/w unless the product sits within
    // rounding distance of a cell boundary, where the reference expression decides
    const double t = pos[i] * P.kn_w[i];
    double fl = floor(t);
    if (t - fl < 1e-9 * (fabs(t) + 1.0) || fl + 1.0 - t < 1e-9 * (fabs(t) + 1.0))
      fl = floor((pos[i] * P.kn[i]) / P.w[i]);
    int k = (int)fl;
    if (pos[i] < 0.0 || pos[i] >= P.w[i])
    {
      if (P.bc[i] == MTB_BC_CUT)
        return -2;
      if (P.bc[i] == MTB_BC_INF)
        return -1;
      // k mod kn in floating point (all values are integers far below 2^53: exact; an integer
      // modulo by a run-time divisor is a 25-instruction sequence and this branch is the common case
      // for ions that left the periodic box)
      const double kn = (double)P.kn[i];
      double r = fl - kn * floor(fl * P.inv_kn[i]);
      if (r < 0.0)
        r += kn;
      else if (r >= kn)
        r -= kn;
      k = (int)r;
    }
    kc[i] = k;
  }
  {
    // distance map (mtb_tables.h): nothing to find around this cell in almost every lookup, and in a
    // periodic box nothing within *safe of path length either.
    // A position exactly on the upper face (pos == w under rounding) can index cell kn: scan then.
    const int inside = (kc[0] < P.kn[0]) & (kc[1] < P.kn[1]) & (kc[2] < P.kn[2]) & (kc[0] >= 0) & (kc[1] >= 0) & (kc[2] >= 0);
    if (inside)
    {
      const uint32_t cell = (uint32_t)kc[0] + (uint32_t)P.kn[0] * ((uint32_t)kc[1] + (uint32_t)P.kn[1] * (uint32_t)kc[2]);
#if MTB_DEVICE_CODE
      const uint32_t n = __ldg(P.cl_dist + cell);
      const float bound = P.cl_safe ? __ldg(P.cl_safe + cell) : 0.0f;
#else
      const uint32_t n = P.cl_dist[cell];
      const float bound = P.cl_safe ? P.cl_safe[cell] : 0.0f;
#endif
      // cell map: nothing within (n - 1) cell edges of path; surface map: every point of this cell is at least
      // `bound` away from any cluster (no scan needed even where a scan could see one)
      if (n || bound > 0.0f)
      {
        *safe = fmax2(n ? (float)(n - 1u) * P.cl_safe_unit : 0.0f, bound);
        return -1;
      }
    }
  }
  // the scan range is only needed here: almost every look-up has returned through the distance map above
  for (int i = 0; i < 3; ++i)
  {
    k1[i] = kc[i] - P.cl_ks[i];
    k2[i] = kc[i] + P.cl_ks[i];
    if (k1[i] < 0 && P.bc[i] != MTB_BC_PBC)
      k1[i] = 0;
    if (k2[i] >= P.kn[i] && P.bc[i] != MTB_BC_PBC)
      k2[i] = P.kn[i] - 1;
  }
  for (int j0 = k1[0]; j0 <= k2[0]; ++j0)
  {
    const int c0 = wrap_cell(j0, P.kn[0]);
    for (int j1 = k1[1]; j1 <= k2[1]; ++j1)
    {
      const int c1 = wrap_cell(j1, P.kn[1]);
      for (int j2 = k1[2]; j2 <= k2[2]; ++j2)
      {
        const int c2 = wrap_cell(j2, P.kn[2]);
        int l = P.cl_hash[c0 + P.kn[0] * (c1 + P.kn[1] * c2)];
        while (l >= 0)
        {
          double r2 = 0.0;
          for (int i = 0; i < 3; ++i)
          {
            double dif = pos[i] - P.cl_xyzr[4 * l + i];
            if (P.bc[i] == MTB_BC_PBC)
              dif -= rint(dif / P.w[i]) * P.w[i]; // ::round differs only at exact .5
            r2 += dif * dif;
          }
          const double rr = P.cl_xyzr[4 * l + 3];
          if (r2 < rr * rr)
            return l;
          l = P.cl_next[l];
        }
      }
    }
  }
  return -1;
}

template <class TR>
MTB_HD int
lookup_material(const LaunchParams & P, const BlockCtx & S, double px, double py, double pz, int * cluster, float * safe)
{
  *cluster = -1;
  *safe = 0.0f;
  if (TR::has(F_CLUSTERS) && P.geom_kind == MTB_GEOM_CLUSTERS) // sample_clusters.C:43-55
  {
    const int l = lookup_cluster(P, px, py, pz, safe);
    if (l == -2)
      return -1;
    if (l == -1)
      return 0;
    *cluster = l;
    return 1;
  }
  if (TR::has(F_GEOM_ANY))
  {
    if (P.geom_kind == MTB_GEOM_WIRE) // sample_wire.C:37-46
    {
      const double x = (px / P.w[0]) * 2.0 - 1.0;
      const double y = (py / P.w[1]) * 2.0 - 1.0;
      return (x * x + y * y) > 1.0 ? -1 : 0;
    }
    if (P.geom_kind == MTB_GEOM_BURIED_WIRE) // sample_burried_wire.C:38-55
    {
      if (pz < 0.0 && pz >= -250.0)
        return 1;
      if (pz > P.w[2] || pz < -250.0)
        return -1;
      const double x = (px / P.w[0]) * 2.0 - 1.0;
      const double y = (py / P.w[1]) * 2.0 - 1.0;
      return (x * x + y * y) > 1.0 ? 1 : 0;
    }
  }
  // sample_solid.C:25-29, sample_layers.C:26-49: the first layer whose cumulative thickness exceeds x;
  // beyond the stack -> last layer
  // (a stack of layers of one and the same material, e.g. inputs/samplelayers_zro2_multilayer.in, needs no search)
  if (P.geom_kind == MTB_GEOM_SOLID || P.n_layers == 1 || P.one_material)
    return 0;
  int lo = 0, hi = P.n_layers - 1;
  while (lo < hi)
  {
    const int mid = (lo + hi) >> 1;
    if (px < S.layer_cum[mid])
      hi = mid;
    else
      lo = mid + 1;
  }
  return S.layer_mat[lo];
}

// ---------------------------------------------------------------------------------------------
// lane state
// ---------------------------------------------------------------------------------------------
struct Lane
{
  // ion in flight
  double px, py, pz, E;
  float dx, dy, dz;
  float Ecur;      // float copy of E (saves double->float conversions on the SFU pipe)
  uint32_t ic;
  uint64_t uid;
  uint32_t packed;
  int32_t tag;
  int32_t pcls;    // projectile class of the ion in flight, -1: this lane's per-primary rows
  float dsafe;     // clusters geometry: path length left before the next lookup can matter (cl_dist)
  // current cascade
  uint32_t prim;   // index of the cascade's primary within this launch (global = first_index + prim)
  int32_t prim_pcls;
  int32_t pZ;
  float pm, Ef;
  double casEel, casEnuc;
  uint32_t casVac, casRepl, casSteps, casIons;
};

// Select the projectile class after L.packed changed (pop, hand-over to a recoil).
MTB_HD void
set_species(Lane & L, const BlockCtx &)
{
  const uint32_t species = L.packed & SPECIES_MASK;
  // a popped ion waits for this value behind a global load: no dependent table look-up here
  L.pcls = species == SPECIES_PRIMARY ? L.prim_pcls : (int32_t)(species - SPECIES_CLASS0);
}

// Projectile class of a primary: a target class, a registered primary species, or -1.
// (Which primary species happen to be registered as classes depends on what the handle has run before; that must not
// change a result — mytrim_uo2 deals chunks of fission events over several GPUs — so the rows of a registered primary
// species are produced by the very functions a lane uses for its private rows: primary_class_rows() below.)
template <class TR>
MTB_HD int
find_class(const LaunchParams & P, const BlockCtx & S, int Z, float m)
{
  const int n = P.n_pclass;
  for (int c = 0; c < n; ++c)
    if (S.pclass[c].Z == Z && S.pclass[c].m == m)
      return c;
  return -1;
}

MTB_HD float4_t
as_row(const PairM & v)
{
  float4_t r = {v.a, v.K, v.C2, v.sk};
  return r;
}
MTB_HD float4_t
as_row(const PairE & v)
{
  float4_t r = {v.my, v.ec, v.inv_ai, v.sfi};
  return r;
}

// A primary whose (Z, m) has no class (e.g. a fission fragment with its own mass) gets private rows
// [ProjClass | PairM per material | PairE per target class] in the lane's scratch area.
#if defined(MTB_ROWS_NOINLINE) && MTB_DEVICE_CODE
__device__ __noinline__ void
#else
MTB_HD_COLD void
#endif
build_custom_rows(const LaunchParams & P, const BlockCtx & S, float4_t * rows, int Z, float m)
{
  const ProjClass c = make_proj_class(S.ionz[Z], Z, m);
  float4_t r0 = {c.m2, c.inv_km, c.m, c.fz};
  float4_t r1 = {c.z023, c.cbrt, c.lfctr, 0.0f};
#if MTB_DEVICE_CODE
  r1.w = __int_as_float(c.Z);
#else
  {
    union { int32_t i; float f; } cv;
    cv.i = c.Z;
    r1.w = cv.f;
  }
#endif
  rows[0] = r0;
  rows[1] = r1;
  for (int mi = 0; mi < P.n_materials; ++mi)
    rows[2 + mi] = as_row(make_pair_m(c, S.materials[mi], P.tmin));
  for (int tc = 0; tc < P.n_tclass; ++tc)
    rows[2 + P.n_materials + tc] = as_row(make_pair_e(c, S.elements[P.tclass_elem[tc]]));
}

// Table rows of projectile class `pc` when it is a registered PRIMARY species (pc >= n_tclass): the same arithmetic as
// build_custom_rows, so that a primary follows the same trajectory whether its species got a class or not.  Runs on
// the device for the engine (mtb_engine.cu: primary_class_rows_kernel) and on the host for the host build of the loop.
MTB_HD void
primary_class_rows(int pc, int n_materials, int n_tclass, float tmin, const DevIonZ * ionz, const DevMaterial * materials,
                   const DevElement * elements, const int32_t * tclass_elem, ProjClass * pclass, PairM * pairm, PairE * paire)
{
  const ProjClass c = make_proj_class(ionz[pclass[pc].Z], pclass[pc].Z, pclass[pc].m);
  pclass[pc] = c;
  for (int mi = 0; mi < n_materials; ++mi)
    pairm[pc * n_materials + mi] = make_pair_m(c, materials[mi], tmin);
  for (int tc = 0; tc < n_tclass; ++tc)
    paire[pc * n_tclass + tc] = make_pair_e(c, elements[tclass_elem[tc]]);
}

MTB_HD ProjClass
row_class(const float4_t * rows)
{
  const float4_t r0 = rows[0], r1 = rows[1];
  ProjClass c;
  c.m2 = r0.x;
  c.inv_km = r0.y;
  c.m = r0.z;
  c.fz = r0.w;
  c.z023 = r1.x;
  c.cbrt = r1.y;
  c.lfctr = r1.z;
#if MTB_DEVICE_CODE
  c.Z = __float_as_int(r1.w);
#else
  {
    union { int32_t i; float f; } cv;
    cv.f = r1.w;
    c.Z = cv.i;
  }
#endif
  return c;
}

MTB_HD int
current_Z(const Lane & L, const BlockCtx & S, const float4_t * rows)
{
  return (L.pcls >= 0 || !rows) ? S.pclass[L.pcls].Z : row_class(rows).Z;
}

// Clusters geometry: what is left of Lane::dsafe travels with a suspended ion (rounded DOWN to whole units, so only
// look-ups that would have returned the matrix are skipped).  Without it every popped or adopted ion started with a
// 27-cell look-up: 19 % of the warp instructions of the tests/uo2 workload at 3 active lanes (profiles/r02_ncu_uo2.md).
MTB_HD uint32_t
safe_bits_of(const LaunchParams & P, float dsafe)
{
  const float q = fmin2(fmax2(dsafe, 0.0f) * P.cl_inv_safe_unit, (float)SAFE_MAX);
  return (uint32_t)(int)q << SAFE_SHIFT;
}

MTB_HD void
stack_store(StackEntry * dst, const Lane & L, uint32_t safe_bits = 0u)
{
#if MTB_DEVICE_CODE
  // 8-byte stores for the doubles (they sit in aligned register pairs already: a 16-byte store would
  // need four moves to line a quad up), 16-byte stores for the rest; no local-memory staging
  double * dd = reinterpret_cast<double *>(dst);
  dd[0] = L.px;
  dd[1] = L.py;
  dd[2] = L.pz;
  dd[3] = L.E;
  uint4 * d = reinterpret_cast<uint4 *>(dst);
  d[2] = make_uint4(__float_as_uint(L.dx), __float_as_uint(L.dy), __float_as_uint(L.dz), L.ic);
  d[3] = make_uint4((uint32_t)L.uid, (uint32_t)(L.uid >> 32), L.packed | safe_bits, (uint32_t)L.tag);
#else
  StackEntry e;
  e.pos[0] = L.px;
  e.pos[1] = L.py;
  e.pos[2] = L.pz;
  e.E = L.E;
  e.dir[0] = L.dx;
  e.dir[1] = L.dy;
  e.dir[2] = L.dz;
  e.ic = L.ic;
  e.uid = L.uid;
  e.packed = L.packed | safe_bits;
  e.tag = L.tag;
  *dst = e;
#endif
}

MTB_HD void
stack_load(const StackEntry * src, Lane & L)
{
#if MTB_DEVICE_CODE
  const uint4 * s = reinterpret_cast<const uint4 *>(src);
  const uint4 a = s[0], b = s[1], c = s[2], d = s[3];
  L.px = __longlong_as_double((long long)(((unsigned long long)a.y << 32) | a.x));
  L.py = __longlong_as_double((long long)(((unsigned long long)a.w << 32) | a.z));
  L.pz = __longlong_as_double((long long)(((unsigned long long)b.y << 32) | b.x));
  L.E = __longlong_as_double((long long)(((unsigned long long)b.w << 32) | b.z));
  L.Ecur = (float)L.E;
  L.dsafe = 0.0f;
  L.dx = __uint_as_float(c.x);
  L.dy = __uint_as_float(c.y);
  L.dz = __uint_as_float(c.z);
  L.ic = c.w;
  L.uid = ((uint64_t)d.y << 32) | d.x;
  L.packed = d.z; // (the caller strips the safe-distance bits of the clusters geometry: restore_safe)
  L.tag = (int32_t)d.w;
#else
  const StackEntry e = *src;
  L.px = e.pos[0];
  L.py = e.pos[1];
  L.pz = e.pos[2];
  L.E = e.E;
  L.Ecur = (float)e.E;
  L.dsafe = 0.0f;
  L.dx = e.dir[0];
  L.dy = e.dir[1];
  L.dz = e.dir[2];
  L.ic = e.ic;
  L.uid = e.uid;
  L.packed = e.packed;
  L.tag = e.tag;
#endif
}

// after a pop / an adoption in the clusters geometry: Lane::dsafe from the bits the suspended ion carried
template <class TR>
MTB_HD void
restore_safe(const LaunchParams & P, Lane & L)
{
  if (TR::has(F_CLUSTERS))
  {
    L.dsafe = (float)(L.packed >> SAFE_SHIFT) * P.cl_safe_unit;
    L.packed &= ~((uint32_t)SAFE_MAX << SAFE_SHIFT);
  }
}

// One half of an ion-log entry (birth: state = -1, position/energy at birth; death: final state,
// position and energy); mtb_get_ion_log joins the halves by uid.
MTB_HD_COLD void
ionlog_append(const LaunchParams & P, double x, double y, double z, double E, uint64_t uid, uint32_t prim, int Z, uint32_t packed,
              int32_t tag, int state)
{
  const unsigned long long i = MTB_ATOMIC_ADD(&P.u64[CNT_IONLOG_N], 1ull);
  if (i >= P.ionlog_cap)
    return;
  mtb_ion_log & o = P.ionlog[i];
  const bool birth = state < 0;
  o.pos0[0] = birth ? x : 0.0;
  o.pos0[1] = birth ? y : 0.0;
  o.pos0[2] = birth ? z : 0.0;
  o.pos1[0] = birth ? 0.0 : x;
  o.pos1[1] = birth ? 0.0 : y;
  o.pos1[2] = birth ? 0.0 : z;
  o.E0 = birth ? E : 0.0;
  o.E1 = birth ? 0.0 : E;
  o.uid = uid;
  o.primary = P.first_index + prim;
  o.Z = Z;
  o.gen = (int32_t)((packed >> GEN_SHIFT) & GEN_MASK);
  o.tag = tag;
  o.state = state;
}

template <class TR>
MTB_HD void
log_birth(const LaunchParams & P, const Lane & L, int Z)
{
  if (!tally_on<TR>(P, MTB_TALLY_IONLOG))
    return;
  if (P.ionlog_z && Z != P.ionlog_z)
    return;
  ionlog_append(P, L.px, L.py, L.pz, L.E, L.uid, L.prim, Z, L.packed, L.tag, -1);
}

// an ion has stopped (or left the sample) at (x, y, z): primary record + death half of the ion log
template <class TR>
MTB_HD void
finish_ion(const LaunchParams & P, const BlockCtx & S, const Lane & L, const float4_t * rows, int state, double x, double y,
           double z)
{
#ifndef MTB_NO_RECORDS
#define MTB_NO_RECORDS 0 // experiment: per-primary records compiled out (profiles/r02_variant_sweeps.md)
#endif
  if (!MTB_NO_RECORDS && !TR::has(F_NOREC) && (L.packed & FLAG_PRIMARY)) // set only when records were asked for (one test on every exit of an ion)
  {
    mtb_record & r = P.records[L.prim];
    r.pos[0] = x;
    r.pos[1] = y;
    r.pos[2] = z;
    r.E = L.E;
    r.state = state;
    r.primary_steps = L.ic;
  }
  if (tally_on<TR>(P, MTB_TALLY_IONLOG))
  {
    const int Z = current_Z(L, S, rows);
    if (P.ionlog_z && Z != P.ionlog_z)
      return;
    ionlog_append(P, x, y, z, L.E, L.uid, L.prim, Z, L.packed, L.tag, state);
  }
}

// `repl` selects the replacement histogram (it follows the vacancy histogram in shared memory and in P.u64)
MTB_HD void
depth_tally(const LaunchParams & P, const BlockCtx & S, bool repl, int x)
{
  // smem_hist_bins <= hist_bins: the common case needs one (unsigned) test that also covers x < 0 —
  // every instruction on this path is issued for 2-4 lanes in almost every iteration of the warp
  if ((unsigned int)x < (unsigned int)P.smem_hist_bins)
  {
    MTB_ATOMIC_ADD(&S.hist_vac[(unsigned int)x + (repl ? (unsigned int)P.smem_hist_bins : 0u)], 1u);
    return;
  }
  if (x < 0)
    return;
  if (x >= P.hist_bins)
  {
    block_add(&S.blk_u64[CNT_CLAMPED], 1u);
    x = P.hist_bins - 1;
  }
  if (x < P.smem_hist_bins)
    MTB_ATOMIC_ADD(&(repl ? S.hist_repl : S.hist_vac)[x], 1u);
  else
    MTB_ATOMIC_ADD(&P.u64[(repl ? off_repl(P) : off_vac(P)) + (size_t)x], 1ull);
}

// vacancyCreation() of the in-tree subclasses
template <class TR>
MTB_HD void
vacancy_creation(const LaunchParams & P, const BlockCtx & S, Lane & L, const DevMaterial & M,
                 const DevElement & el, double rx, double ry, float Erec, int rec_gen)
{
  switch (TR::has(F_VACMODEL) ? P.vacancy_model : (int)MTB_VAC_COUNT)
  {
    case MTB_VAC_COUNT: // trim.C:439-443
      L.casVac++;
      break;
    case MTB_VAC_NRT: // apps/src/TrimRange.C:31-47
    {
      const float ed = 0.0115f * fpow(el.fz, -7.0f / 3.0f) * Erec;
      const float kd = 0.1337f * fpow(el.fz, 2.0f / 3.0f) * frsqrt(el.m);
      const float g = 3.4008f * fpow(ed, 1.0f / 6.0f) + 0.40244f * fpow(ed, 0.75f) + ed;
      const float Ev = fdiv(Erec, 1.0f + kd * g);
      if (Ev >= el.Edisp)
        L.casVac += (Ev >= el.Edisp * 2.5f) ? (uint32_t)(Ev * 0.4f / el.Edisp) : 1u;
      break;
    }
    case MTB_VAC_KP: // TrimPrimaries::vacancyCreation — trim.C:445-464
      L.casVac++;
      if (rec_gen == P.follow_max_gen)
      {
        const float ed = 0.0115f * fpow(M.az, -7.0f / 3.0f) * Erec;
        const float g = 3.4008f * fpow(ed, 1.0f / 6.0f) + 0.40244f * fpow(ed, 0.75f) + ed;
        const float kd = 0.1337f * fpow(M.az, 2.0f / 3.0f) * frsqrt(M.am);
        const float Ev = fdiv(Erec, 1.0f + kd * g);
        L.casVac += (uint32_t)(int)(0.8f * Ev / (2.0f * el.Edisp));
      }
      break;
    default:
      break;
  }
  const int x = (int)rx; // truncation toward zero — TrimVacCount.C:35 (the depth histogram itself: caller)
  if (tally_on<TR>(P, MTB_TALLY_VAC_ENERGY) && x >= 0) // TrimVacEnergyCount.C:31-53
  {
    int le = (int)flog(Erec);
    le = le < 0 ? 0 : (le >= P.evac_rows ? P.evac_rows - 1 : le);
    int xb = x;
    if (xb >= P.hist_bins)
    {
      block_add(&S.blk_u64[CNT_CLAMPED], 1u);
      xb = P.hist_bins - 1;
    }
    MTB_ATOMIC_ADD(&P.u64[off_evac(P) + (size_t)le * (size_t)P.hist_bins + (size_t)xb], 1ull);
  }
  if (tally_on<TR>(P, MTB_TALLY_VACMAP)) // TrimVacMap::vacancyCreation — trim.C:483-501
  {
    int vx = (int)((rx * MTB_VMAP_NX) / P.w[0]);
    int vy = (int)((ry * MTB_VMAP_NY) / P.w[1]);
    vx -= (vx / MTB_VMAP_NX) * MTB_VMAP_NX;
    vy -= (vy / MTB_VMAP_NY) * MTB_VMAP_NY;
    int s = -1;
    if (el.Z == P.vmap_z[0])
      s = 0;
    else if (el.Z == P.vmap_z[1])
      s = 1;
    else if (el.Z == P.vmap_z[2])
      s = 2;
    if (s >= 0 && vx >= 0 && vy >= 0)
      MTB_ATOMIC_ADD(&P.u64[off_vmap(P) + (size_t)((vx * MTB_VMAP_NY + vy) * 3 + s)], 1ull);
  }
}

// close a subtree of a cascade (the whole cascade unless lanes shared it): per-primary record and
// block totals.  Record counters are accumulated atomically because several lanes may have worked
// on the same primary; records are zeroed before the launch.
template <class TR>
#if defined(MTB_CLOSE_NOINLINE) && MTB_DEVICE_CODE
__device__ __noinline__ void
#else
MTB_HD void
#endif
close_subtree(const LaunchParams & P, const BlockCtx & S, Lane & L, uint32_t n_prim)
{
  if (!TR::has(F_NOREC) && P.records)
  {
    mtb_record & r = P.records[L.prim];
    MTB_ATOMIC_ADD(&r.Eel, L.casEel);
    MTB_ATOMIC_ADD(&r.Enuc, L.casEnuc);
    MTB_ATOMIC_ADD(&r.vacancies, L.casVac);
    MTB_ATOMIC_ADD(&r.replacements, L.casRepl);
    MTB_ATOMIC_ADD(&r.steps, L.casSteps);
    MTB_ATOMIC_ADD(&r.ions, L.casIons);
  }
  block_add(&S.blk_u64[CNT_VAC], L.casVac);
  block_add(&S.blk_u64[CNT_REPL], L.casRepl);
  block_add(&S.blk_u64[CNT_STEPS], L.casSteps);
  block_add(&S.blk_u64[CNT_IONS], L.casIons);
  block_add(&S.blk_u64[CNT_QUEUED], L.casIons - n_prim);
  block_add(&S.blk_u64[CNT_PRIMARIES], n_prim);
  MTB_ATOMIC_ADD(&S.blk_f64[0], L.casEel);
  MTB_ATOMIC_ADD(&S.blk_f64[1], L.casEnuc);
}

// ---------------------------------------------------------------------------------------------
// work-sharing pool (device only): one bounded MPMC ring (after D. Vyukov) per CTA in shared memory.
// When a launch has fewer primaries than lanes, lanes without work adopt suspended ions that busy
// lanes of the same CTA donate instead of pushing them on their private stacks; a heavy cascade
// fans out over the CTA within a few generations.  Both operations are non-blocking attempts.
// (A single device-wide ring was measured first: 1e5 polling lanes on one L2 line made the kernel
// 50x slower.  Shared memory has no such problem.)
// ---------------------------------------------------------------------------------------------
#if MTB_DEVICE_CODE
#ifdef MTB_POOL_NOINLINE
#define MTB_POOL_FN __device__ __noinline__
#else
#define MTB_POOL_FN __device__ __forceinline__
#endif
MTB_D unsigned long long
vload(const unsigned long long * p)
{
  return *reinterpret_cast<const volatile unsigned long long *>(p);
}

// sequence numbers of the ring slots carry the hand-over of the payload: acquire / release at CTA scope
MTB_D unsigned long long
load_acquire(const unsigned long long * p)
{
  unsigned long long v;
  asm volatile("ld.acquire.cta.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

MTB_D void
store_release(unsigned long long * p, unsigned long long v)
{
  asm volatile("st.release.cta.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Claim a ring slot for writing (which = POOL_ENQ) or reading (which = POOL_DEQ): the ticket dance of the
// bounded MPMC queue.  Inlined: as a real call (-DMTB_POOL_NOINLINE) the tests/uo2 workload was measured 9 % slower
// in round 2 (profiles/r02_variant_sweeps.md), although the inlined form costs the sharing kernels registers.
MTB_POOL_FN PoolSlot *
pool_claim(PoolSlot * pool, unsigned long long * ctl, int which, unsigned long long * ticket)
{
  const unsigned long long mask = MTB_POOL_SLOTS - 1;
  const unsigned long long want = which == POOL_ENQ ? 0ull : 1ull; // slot sequence relative to the ticket
  unsigned long long pos = vload(&ctl[which]);
  for (int tries = 0; tries < 4; ++tries)
  {
    PoolSlot * slot = pool + (pos & mask);
    const long long dif = (long long)(load_acquire(&slot->seq) - (pos + want));
    if (dif == 0)
    {
      const unsigned long long seen = atomicCAS(&ctl[which], pos, pos + 1);
      if (seen == pos)
      {
        if (which == POOL_ENQ)
          atomicAdd(&ctl[POOL_WORKING], 1ull); // the entry counts as outstanding work from now on
        *ticket = pos;
        return slot;
      }
      pos = seen;
    }
    else if (dif < 0)
      return nullptr; // full (push) / empty (pop)
    else
      pos = vload(&ctl[which]);
  }
  return nullptr;
}

MTB_D bool
pool_try_push(const BlockCtx & S, const Lane & ion, uint64_t prim, uint32_t safe_bits)
{
  unsigned long long pos;
  PoolSlot * slot = pool_claim(S.pool, S.pool_ctl, POOL_ENQ, &pos);
  if (!slot)
    return false;
  slot->prim = prim;
  stack_store(&slot->e, ion, safe_bits);
  store_release(&slot->seq, pos + 1); // publishes the payload
  return true;
}

MTB_D bool
pool_try_pop(const BlockCtx & S, Lane & ion, uint64_t * prim)
{
  unsigned long long pos;
  PoolSlot * slot = pool_claim(S.pool, S.pool_ctl, POOL_DEQ, &pos);
  if (!slot)
    return false;
  *prim = slot->prim;
  stack_load(&slot->e, ion);
  store_release(&slot->seq, pos + MTB_POOL_SLOTS); // hands the slot back to the producers
  return true;
}
#endif

// Stack cursor of a lane: the byte offset, within P.stacks, of its next free entry.  A lane owns
// MTB_STACK_DEPTH consecutive 64-byte entries, so bits 6.. of the cursor hold lane * MTB_STACK_DEPTH +
// depth: the depth needs no register of its own and an entry address is one 32 x 64-bit add.
// (Depth MTB_STACK_DEPTH itself would carry into the lane index: a stack is full at MTB_STACK_DEPTH - 1.)
static_assert(sizeof(StackEntry) == 64 && (MTB_STACK_DEPTH & (MTB_STACK_DEPTH - 1)) == 0, "stack cursor encoding");
#define MTB_STACK_DEPTH_BITS ((uint32_t)(MTB_STACK_DEPTH - 1) << 6)

MTB_HD StackEntry *
stack_entry(const LaunchParams & P, uint32_t cursor)
{
  return reinterpret_cast<StackEntry *>(reinterpret_cast<unsigned char *>(P.stacks) + cursor);
}

MTB_HD int
stack_depth(uint32_t cursor)
{
  return (int)((cursor & MTB_STACK_DEPTH_BITS) >> 6);
}

// A primary the tables cannot describe (out of line: the test runs once per cascade and must not disturb the
// register allocation and the layout of the collision loop).
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline
#endif
bool
primary_is_invalid(const mtb_ion & src)
{
  const double d2 = src.dir[0] * src.dir[0] + src.dir[1] * src.dir[1] + src.dir[2] * src.dir[2];
  return (unsigned int)(src.Z - 1) >= (unsigned int)MTB_NZ || !(src.m > 0.0 && src.m < 1.0e6) ||
         !(src.E >= 0.0 && src.E <= 1.0e15) || !(d2 > 0.0 && d2 < 1.0e300);
}

// Suspend an ion: on the lane's private stack, or — when lanes are idle — in the shared pool.
template <class TR>
MTB_HD void
suspend_ion(const LaunchParams & P, const BlockCtx & S, uint32_t & sp, const Lane & ion, uint64_t prim, float e_small)
{
#if MTB_DEVICE_CODE
  // Donate only when BOTH ions of the pair carry a subtree worth a hand-over (e_small = the smaller of
  // the two energies): a lane that gives its big ion away and keeps a 30 eV recoil is idle itself
  // three steps later, and every adoption costs a few hundred instructions at one active lane.
  const uint32_t safe_bits = TR::has(F_CLUSTERS) ? safe_bits_of(P, ion.dsafe) : 0u;
  if (TR::kShare && e_small >= P.share_min_E && vload(&S.pool_ctl[POOL_IDLE]) > 0 && pool_try_push(S, ion, prim, safe_bits))
    return;
#else
  const uint32_t safe_bits = TR::has(F_CLUSTERS) ? safe_bits_of(P, ion.dsafe) : 0u;
#endif
  (void)S;
  (void)prim;
  if ((sp & MTB_STACK_DEPTH_BITS) != MTB_STACK_DEPTH_BITS)
  {
    stack_store(stack_entry(P, sp), ion, safe_bits);
    sp += (uint32_t)sizeof(StackEntry);
  }
  else
    MTB_ATOMIC_ADD(&P.u64[CNT_ERROR], 1ull);
}

// ---------------------------------------------------------------------------------------------
// the lane loop.  EVENTS=true is the single-ion mode behind mtb_trim_one: one ion, recoils are
// never followed, every collision is reported.
// ---------------------------------------------------------------------------------------------
template <class TR>
MTB_HD void
lane_loop(const LaunchParams & P, const BlockCtx & S, uint32_t lane_global)
{
  constexpr bool EVENTS = TR::kEvents;
  const int potential = TR::has(F_POTENTIAL) ? P.potential : (int)MTB_POT_UNIVERSAL;
  Lane L;
  float4_t * const rows =
      TR::kCustom ? P.custom_rows + (size_t)lane_global * (size_t)(2 + P.n_materials + P.n_tclass) : nullptr;
  uint32_t sp = lane_global * (uint32_t)(MTB_STACK_DEPTH * sizeof(StackEntry)); // stack cursor, see stack_entry()
  int sp_max = 0;
  bool active = false, open = false, started = false, done = false, no_more = false, idle = false;
  uint32_t cas_prim = 0;
  unsigned long long idle_polls = 0;
  unsigned long long n_events = 0;
  L.prim = 0;
  L.casEel = L.casEnuc = 0.0;
  L.casVac = L.casRepl = L.casSteps = L.casIons = 0;

  // Loop structure: [refill] -> [warp vote] -> [one collision].  No lane leaves the loop before the
  // whole warp is out of work, and no `continue` jumps back to the loop head from divergent code, so
  // the 32 lanes reconverge at the vote in every iteration (a lane-level early exit or continue lets
  // the compiler split the warp for good — measured: 15 instead of 29 active threads).
#ifndef MTB_DEFER
#define MTB_DEFER 0
#endif
#ifndef MTB_DEFER_EVERY
#define MTB_DEFER_EVERY 2
#endif
  // Batched hand-over (experiment, -DMTB_DEFER=<T>): the end-of-step paths (recoil hand-over, stack push/pop, ion end,
  // refill) cost ~25 % of the warp instructions at 2-9 active lanes because almost every iteration has SOME lane that
  // needs them.  With MTB_DEFER a lane that needs one parks its collision result in registers and sits out until at
  // least T lanes of the warp need the hand-over phase or MTB_DEFER_EVERY iterations have passed; the phase then runs
  // once for all of them.
  constexpr bool DEFER = MTB_DEFER > 0 && !EVENTS && !TR::kShare && MTB_DEVICE_CODE;
  uint32_t pend = 0; // bit 0: a collision result waits for its hand-over; bit 1: recoil to follow; bits 2..: state
  float sqx = 0.f, sqy = 0.f, sqz = 0.f, sErec = 0.f, smx = 0.f, smy = 0.f, smz = 0.f, sdsafe = 0.f;
  uint32_t sw3 = 0, srp = 0;
  int32_t smtag = 0;
  uint32_t iter = 0;
#ifndef MTB_POLL_EVERY
#define MTB_POLL_EVERY 1 // power of two
#endif
  uint32_t trip = 0;

  for (;;)
  {
    if (TR::kShare && MTB_POLL_EVERY > 1)
      ++trip;
    bool resolve = true;
#if MTB_DEVICE_CODE
    if (DEFER)
    {
      const unsigned int need = __ballot_sync(0xffffffffu, pend != 0 || (!active && !done));
      ++iter;
      resolve = __popc(need) >= MTB_DEFER || (iter % MTB_DEFER_EVERY) == 0 || need == 0xffffffffu;
    }
#endif
    if (DEFER && resolve && pend)
    {
      // ---------------- deferred "who flies next" (same logic as at the end of the collision below) ----------------
      const int state = (int)(pend >> 2);
      const bool follow = (pend & 2u) != 0;
      pend = 0;
      const double mvx = (double)smx, mvy = (double)smy, mvz = (double)smz;
      const float E2 = L.Ecur;
      if (follow)
      {
        L.casIons++;
        const float qs = frsqrt(sqx * sqx + sqy * sqy + sqz * sqz);
        const uint64_t ruid = child_uid(L.uid, L.ic, sw3);
        const bool keep_projectile = (state == MTB_MOVING) && (E2 <= sErec);
        if (state == MTB_MOVING && !keep_projectile)
        {
          Lane T = L;
          T.px = L.px + mvx;
          T.py = L.py + mvy;
          T.pz = L.pz + mvz;
          suspend_ion<TR>(P, S, sp, T, L.prim, sErec);
        }
        if (state != MTB_MOVING)
          finish_ion<TR>(P, S, L, rows, state, L.px + mvx, L.py + mvy, L.pz + mvz);
        if (keep_projectile)
        {
          Lane R;
          R.px = L.px; R.py = L.py; R.pz = L.pz;
          R.E = (double)sErec;
          R.dx = sqx * qs; R.dy = sqy * qs; R.dz = sqz * qs;
          R.ic = 0;
          R.uid = ruid;
          R.packed = srp;
          R.tag = smtag;
          R.dsafe = sdsafe;
          if (tally_on<TR>(P, MTB_TALLY_IONLOG))
          {
            R.prim = L.prim;
            log_birth<TR>(P, R, S.pclass[(srp & SPECIES_MASK) - SPECIES_CLASS0].Z);
          }
          suspend_ion<TR>(P, S, sp, R, L.prim, E2);
          L.px += mvx;
          L.py += mvy;
          L.pz += mvz;
        }
        else
        {
          L.E = (double)sErec;
          L.Ecur = sErec;
          L.dx = sqx * qs; L.dy = sqy * qs; L.dz = sqz * qs;
          L.ic = 0;
          L.uid = ruid;
          L.packed = srp;
          L.tag = smtag;
          L.pcls = (int32_t)((srp & SPECIES_MASK) - SPECIES_CLASS0);
          L.dsafe = sdsafe;
          log_birth<TR>(P, L, S.pclass[L.pcls].Z);
        }
      }
      else
      {
        finish_ion<TR>(P, S, L, rows, state, L.px + mvx, L.py + mvy, L.pz + mvz);
        active = false;
      }
    }
    // ---------------- refill: next suspended ion, else next primary ----------------
    if (resolve && !done && !active)
    {
      if (sp & MTB_STACK_DEPTH_BITS)
      {
        sp -= (uint32_t)sizeof(StackEntry);
        stack_load(stack_entry(P, sp), L);
        restore_safe<TR>(P, L);
        set_species(L, S);
        active = true;
      }
      else
      {
        if (open)
        {
          close_subtree<TR>(P, S, L, TR::kShare ? cas_prim : 1u);
          open = false;
        }
#ifdef MTB_REFILL_UNROLL1
#pragma unroll 1
#endif
        while (!no_more)
        {
          unsigned long long idx;
          if (EVENTS)
          {
            // event mode: lane i follows ion i of the batch (mtb_trim_one: a batch of one)
            if (started || lane_global >= P.n_primaries)
            {
              no_more = true;
              break;
            }
            idx = lane_global;
          }
#if MTB_DEVICE_CODE
          else if (TR::kShare)
          {
            // The first primary of every lane is dealt round-robin over the CTAs (CTA-local counter), so
            // that a launch with fewer primaries than lanes gives every CTA something to share among
            // its lanes; the rest comes from the device-wide counter like in the plain kernels (a
            // static split of ALL primaries over the CTAs cost 10 % at 16 primaries per lane).
            const unsigned long long k = atomicAdd(&S.pool_ctl[POOL_CTL_COUNT], 1ull);
            if (k < (unsigned long long)blockDim.x)
              idx = (unsigned long long)blockIdx.x + (unsigned long long)gridDim.x * k;
            else
              idx = (unsigned long long)gridDim.x * blockDim.x + atomicAdd(&P.u64[CNT_NEXT_PRIMARY], 1ull);
          }
#endif
          else
            idx = MTB_ATOMIC_ADD(&P.u64[CNT_NEXT_PRIMARY], 1ull);
          started = true;
          if (idx >= P.n_primaries)
          {
            no_more = true;
            break;
          }
          if (P.index_list)
            idx = P.index_list[idx];
          const mtb_ion & src = P.primaries ? P.primaries[idx] : P.beam;
          const int src_Z = src.Z;
          const float src_m = (float)src.m;
          {
            // A primary the tables cannot describe (Z outside 1..92, m <= 0, negative or non-finite energy, no
            // direction) would index the Z tables out of bounds or fly as NaN: it is skipped and counted in the
            // upper half of the error word; the run then fails with MTB_EINVAL (mtb_engine.cu: sync_and_check).
            // (out of line: inlined, the same test cost the north-star kernel 1.9 % — 64 more instructions around the
            // collision loop and a different register allocation; as a call 0.25 %, profiles/r02_variant_sweeps.md)
            if (primary_is_invalid(src))
            {
              MTB_ATOMIC_ADD(&P.u64[CNT_ERROR], 1ull << 32);
              continue;
            }
          }
          const int cls = find_class<TR>(P, S, src_Z, src_m);
          if (!TR::kCustom && cls < 0)
          {
            // species without a projectile class: hand the primary to the generic kernel
            P.deferred[MTB_ATOMIC_ADD(&P.u64[CNT_DEFERRED], 1ull)] = (uint32_t)idx;
            continue;
          }
          L.px = src.pos[0];
          L.py = src.pos[1];
          L.pz = src.pos[2];
          L.dx = (float)src.dir[0];
          L.dy = (float)src.dir[1];
          L.dz = (float)src.dir[2];
          {
            // v_norm(dir) of the first step (trim.C:85); later steps only trim the rounding drift
            const float inv = frsqrt(L.dx * L.dx + L.dy * L.dy + L.dz * L.dz);
            L.dx *= inv;
            L.dy *= inv;
            L.dz *= inv;
          }
          L.E = src.E;
          L.Ecur = (float)src.E;
          L.dsafe = 0.0f;
          L.ic = 0;
          L.prim = (uint32_t)idx;
          L.uid = EVENTS ? (P.uid_list ? P.uid_list[idx] : P.single_uid + idx) : P.first_index + idx;
          L.packed = SPECIES_PRIMARY | (((uint32_t)src.gen & GEN_MASK) << GEN_SHIFT) | ((!TR::has(F_NOREC) && P.records) ? (uint32_t)FLAG_PRIMARY : 0u);
          L.tag = src.tag;
          L.pZ = src_Z;
          L.pm = src_m;
          L.Ef = (float)src.Ef;
          L.casEel = L.casEnuc = 0.0;
          L.casVac = L.casRepl = L.casSteps = 0;
          L.casIons = 1;
          cas_prim = 1;
          open = true;
          L.prim_pcls = cls;
          if (TR::kCustom && cls < 0)
            build_custom_rows(P, S, rows, src_Z, src_m);
          L.pcls = cls;
          if (!EVENTS)
            log_birth<TR>(P, L, src_Z);
          active = true;
          break;
        }
        if (!active)
        {
          // out of primaries: either done, or (work sharing) adopt what other lanes donate
#if MTB_DEVICE_CODE
          if (TR::kShare)
          {
            if (!idle)
            {
              idle = true;
              atomicAdd(&S.pool_ctl[POOL_IDLE], 1ull);
              atomicAdd(&S.pool_ctl[POOL_WORKING], (unsigned long long)-1ll);
            }
            uint64_t aprim;
            // (Polling only in every 4th / 16th trip of the warp, -DMTB_POLL_EVERY, to spare the working warp mates
            // the ~40 instructions of the attempt, was measured 5-14 % SLOWER on the tests/uo2 workload and 5-8 % on
            // C->W / Xe->ZrO2: how soon an idle lane picks work up matters more.  profiles/r02_variant_sweeps.md)
            if ((MTB_POLL_EVERY == 1 || (trip & (MTB_POLL_EVERY - 1)) == 0) && pool_try_pop(S, L, &aprim))
            {
              // the entry carried its own count in POOL_WORKING; it now belongs to this lane
              idle = false;
              restore_safe<TR>(P, L);
              atomicAdd(&S.pool_ctl[POOL_IDLE], (unsigned long long)-1ll);
              const mtb_ion & src = P.primaries ? P.primaries[aprim] : P.beam;
              L.prim = (uint32_t)aprim;
              L.Ef = (float)src.Ef;
              L.prim_pcls = 0;
              if ((L.packed & SPECIES_MASK) == SPECIES_PRIMARY)
              {
                // only the primary itself flies as its own species: a recoil subtree needs neither the
                // class of the primary nor (per-primary masses) its private rows
                const int aZ = src.Z;
                const float am = (float)src.m;
                L.prim_pcls = find_class<TR>(P, S, aZ, am);
                if (TR::kCustom && L.prim_pcls < 0)
                  build_custom_rows(P, S, rows, aZ, am);
              }
              set_species(L, S);
              L.casEel = L.casEnuc = 0.0;
              L.casVac = L.casRepl = L.casSteps = L.casIons = 0;
              cas_prim = 0;
              open = true;
              active = true;
            }
            else if (vload(&S.pool_ctl[POOL_WORKING]) == 0)
              done = true;
            else if (++idle_polls > (1ull << 26))
            {
              atomicAdd(&P.u64[CNT_ERROR], 1ull); // watchdog: never spin forever on an accounting bug
              done = true;
            }
          }
          else
#endif
            done = true;
        }
      }
    }

    if (MTB_WARP_ALL(done))
      break;
#if MTB_DEVICE_CODE
    if (TR::kShare && MTB_WARP_ALL(!active))
      __nanosleep(400); // the whole warp is polling: back off
#endif

    // Clusters geometry: when ANY lane of the warp has used up its safe distance, every active lane looks its material
    // up in this trip.  A look-up costs the warp its ~150 instructions whether one lane or thirty-two take part (2 % of
    // the lane-steps of the tests/uo2 workload needed one, so half of all trips paid for it: 15 % of the warp
    // instructions at 3 active lanes); taken together, all lanes leave with a fresh safe distance and the next
    // look-up is tens of trips away.  The extra look-ups return the matrix by construction: no result changes.
    bool refresh = false;
    if (TR::has(F_CLUSTERS) && !TR::has(F_MONO))
      refresh = MTB_WARP_ANY(active && !(L.dsafe > 0.0f));

    if (active && !(DEFER && pend))
      do
      {
    // ---------------- one collision: trim.C:74-424 ----------------
    if (!(L.Ecur > 0.0f))
    {
      // the reference would produce NaNs for a projectile without energy; park it instead
      finish_ion<TR>(P, S, L, rows, MTB_INTERSTITIAL, L.px, L.py, L.pz);
      active = false;
      break;
    }
    ++L.ic;
    int cluster = -1;
    int mi = 0;
    if (TR::has(F_MONO))
      ; // material 0 everywhere (sample_solid.C:25-29; sample_layers.C:26-49 never returns vacuum)
    else if (!(TR::has(F_CLUSTERS) && L.dsafe > 0.0f && !refresh)) // else: clusters geometry, still provably in the matrix
    {
      float safe;
      mi = lookup_material<TR>(P, S, L.px, L.py, L.pz, &cluster, &safe);
      L.dsafe = safe;
    }
    if (!TR::has(F_MONO) && mi < 0)
    {
      // vacuum: the reference breaks out with the state still MOVING (trim.C:80-82)
      block_add(&S.blk_u64[CNT_LEFT], 1u);
      --L.ic;
      finish_ion<TR>(P, S, L, rows, MTB_MOVING, L.px, L.py, L.pz);
      active = false;
      break;
    }
    const DevMaterial & M = S.materials[TR::has(F_MONO) ? 0 : mi];
    const int mtag = (TR::has(F_CLUSTERS) && P.geom_kind == MTB_GEOM_CLUSTERS && mi == 1) ? cluster : M.tag;
    L.casSteps++;

    // v_norm(dir) — trim.C:85.  Every ion enters the loop with a unit direction (primaries are
    // normalised when they are loaded, recoils when they are created) and a rotation keeps the norm
    // up to rounding, so 1/|d| = 1.5 - 0.5 |d|^2 is exact to O(1e-13): no MUFU.RSQ on the step path.
    {
      const float inv = fmaf(-0.5f, L.dx * L.dx + L.dy * L.dy + L.dz * L.dz, 1.5f);
      L.dx *= inv;
      L.dy *= inv;
      L.dz *= inv;
    }

    // the four uniforms of this step: one Philox block
    uint32_t w[4];
    philox4x32_10_rk(L.ic, (uint32_t)L.uid, (uint32_t)(L.uid >> 32), 0u, P.rk, w);
    const float r2 = u01(w[0]);
    float hh = u01(w[1]);
    const float r1 = u01(w[3]);

    const float E0 = L.Ecur;
    const bool custom = TR::kCustom && L.pcls < 0;
    const ProjClass pc = custom ? row_class(rows) : S.pclass[L.pcls];
    const LowStop * const lowrow = S.lowstop + pc.Z * P.n_zslots;

    // free flight path and impact parameter — trim.C:88-94, 143-144 (constants of
    // MaterialBase::average from the (projectile class, material) table)
    PairM pm;
    if (custom)
    {
      const float4_t r = rows[2 + mi];
      pm.a = r.x;
      pm.K = r.y;
      pm.C2 = r.z;
      pm.sk = r.w;
    }
    else
      pm = S.pairm[TR::has(F_MONO) ? L.pcls : L.pcls * P.n_materials + mi];
    float ls;
    const float sqrtE0 = fsqrt(E0);
    const float pmax = flight_from_pair(pm, sqrtE0, &ls);
    if (L.ic == 1)
      ls = r1 * fmin2(ls, P.cw);
    const float p = pmax * fsqrt(r2);

    // target element — trim.C:147-156
    int nn = 0;
    if (!TR::has(F_MONO))
      for (; nn < M.n_elem - 1; ++nn)
      {
        hh -= S.elements[M.first_elem + nn].t;
        if (hh <= 0.0f)
          break;
      }
    const DevElement & el = S.elements[TR::has(F_MONO) ? 0 : M.first_elem + nn];

    // element part of MaterialBase::average — material.C:99-108
    PairE pe;
    if (custom)
    {
      const float4_t r = rows[2 + P.n_materials + el.tcls];
      pe.my = r.x;
      pe.ec = r.y;
      pe.inv_ai = r.z;
      pe.sfi = r.w;
    }
    else
      pe = S.paire[TR::has(F_MONO) ? L.pcls : L.pcls * P.n_tclass + el.tcls];
    const float my = pe.my;

    const float sqe = pe.sfi * sqrtE0; // sqrt(eps), eps = fi E — trim.C:159-160
    const float b = p * pe.inv_ai;

    const float see = TR::has(F_MONO) ? (0.0f + element_stopping(pc, lowrow, el, E0, sqrtE0 * pm.sk) * el.t) * M.arho
                                      : material_stopping(pc, lowrow, M, S.elements, E0, sqrtE0 * pm.sk); // trim.C:166
    const float dee_f = ls * see;

    const Scatter sc = magic_scatter(potential, sqe, b);

    // energy bookkeeping — trim.C:275-296.  The running energy is FP64; the float images used by
    // the physics are derived from it once per step.
    const float den_f = pe.ec * sc.s2 * E0;
    double dee = (double)dee_f;
    float E1 = E0 - dee_f; // energy after the electronic loss, before the collision
    if (dee > L.E)
    {
      dee = L.E;
      E1 = 0.0f;
    }
    L.E -= dee;
    L.casEel += dee;
    const float E1p = fmax2(E1, 0.0f);
    double den = (double)den_f;
    float Erec_den = den_f;
    if (den > L.E)
    {
      den = L.E;
      Erec_den = (float)den;
    }
    L.E -= den;
    const float E2 = (float)L.E;
    L.Ecur = E2;
    // recoil momentum p1 d - p2 d' with p = sqrt(2 m E) (trim.C:292-296, 311, 341).  Only its
    // direction is used for a followed recoil, so the cascade kernels form sqrt(E1/2m) times it:
    // E1 d - sqrt(E1 E2) d' (one square root instead of two); the event mode reports the momentum itself.
    const float p1 = EVENTS ? fsqrt(pc.m2 * E1p) : E1p;
    const float p2 = EVENTS ? fsqrt(pc.m2 * E2) : fsqrt(E1p * E2);

    // The recoil is born at the previous collision site (trim.C:306-310), the projectile moves on by
    // dir * (ls - tau).  L.p* stays the collision site until the end of the step and is advanced in place
    // only on the paths where the projectile flies on: keeping both positions alive across the fate logic
    // cost seven register moves on every exit of the step.
    const float flight = (ls - P.tau) * P.inv_scale;
    const float mfx = L.dx * flight, mfy = L.dy * flight, mfz = L.dz * flight;
    const double mvx = (double)mfx, mvy = (double)mfy, mvz = (double)mfz;
    const float dsafe_here = L.dsafe; // of the collision site: a recoil starts there
    L.dsafe -= fabsf(flight);
#define MTB_AHEAD_X (L.px + mvx)
#define MTB_AHEAD_Y (L.py + mvy)
#define MTB_AHEAD_Z (L.pz + mvz)

    // unit vector perpendicular to dir with uniform azimuth (replaces trim.C:322-333)
    float qx, qy, qz;
    {
      const float sg = copysignf(1.0f, L.dz);
      const float a = -frcp(sg + L.dz);
      const float bb = L.dx * L.dy * a;
      float sphi, cphi;
      unit_circle(w[2], &sphi, &cphi);
      const float ex = cphi * (1.0f + sg * L.dx * L.dx * a) + sphi * bb;
      const float ey = cphi * (sg * bb) + sphi * (sg + L.dy * L.dy * a);
      const float ez = cphi * (-sg * L.dx) + sphi * (-L.dy);

      // lab scattering angle: psi = atan2(st, ct + my) — trim.C:336-341
      const float ct = 1.0f - 2.0f * sc.s2;
      const float st = 2.0f * fsqrt(sc.s2 * sc.c2);
      const float X = ct + my;
      const float h2 = st * st + X * X;
      float cpsi = 1.0f, spsi = 0.0f;
      if (h2 > 1e-30f)
      {
        const float ih = frsqrt(h2);
        cpsi = X * ih;
        spsi = st * ih;
      }
      const float nx = L.dx * cpsi + ex * spsi;
      const float ny = L.dy * cpsi + ey * spsi;
      const float nz = L.dz * cpsi + ez * spsi;
      // recoil momentum = p1*dir_old - p2*dir_new
      qx = fmaf(-p2, nx, L.dx * p1);
      qy = fmaf(-p2, ny, L.dy * p1);
      qz = fmaf(-p2, nz, L.dz * p1);
      L.dx = nx;
      L.dy = ny;
      L.dz = nz;
    }

    // CUT boundaries — trim.C:344-352
    int state = MTB_MOVING;
    if (TR::has(F_CUT) &&
        ((P.bc[0] == MTB_BC_CUT && (MTB_AHEAD_X > P.w[0] || MTB_AHEAD_X < 0.0)) ||
        (P.bc[1] == MTB_BC_CUT && (MTB_AHEAD_Y > P.w[1] || MTB_AHEAD_Y < 0.0)) ||
        (P.bc[2] == MTB_BC_CUT && (MTB_AHEAD_Z > P.w[2] || MTB_AHEAD_Z < 0.0))))
    {
      state = MTB_LOST;
      block_add(&S.blk_u64[CNT_LOST], 1u);
    }

    // fate of recoil and projectile — trim.C:357-411
    const float Erec = Erec_den - el.Elbind;
    const int rec_gen = (int)((L.packed >> GEN_SHIFT) & GEN_MASK) + 1;
    bool above = false, follow = false;
    if (state != MTB_LOST)
    {
      if (Erec > el.Edisp - el.Elbind)
      {
        above = true;
        if (tally_on<TR>(P, MTB_TALLY_PHONON))
          L.casEnuc += (double)el.Elbind; // TrimPhononOut::followRecoil
        follow = !EVENTS && (!TR::has(F_FOLLOW) || P.follow == MTB_FOLLOW_ALL ||
                             (P.follow == MTB_FOLLOW_GEN_LT && rec_gen < P.follow_max_gen));
        const bool vacancy = E2 > el.Edisp;
        if (vacancy)
          vacancy_creation<TR>(P, S, L, M, el, L.px, L.py, Erec, rec_gen);
        else
        {
          L.casRepl++;
          state = (pc.Z == el.Z) ? MTB_REPLACEMENT : MTB_SUBSTITUTIONAL;
        }
        // TrimVacCount::vacancyCreation / replacementCollision (TrimVacCount.C:31-53): one tally site
        if (tally_on<TR>(P, MTB_TALLY_VAC_DEPTH))
          depth_tally(P, S, !vacancy, (int)L.px);
      }
      else
      {
        if (tally_on<TR>(P, MTB_TALLY_RANGE)) // TrimRange::dissipateRecoilEnergy
        {
          const unsigned long long i = MTB_ATOMIC_ADD(&P.u64[CNT_RANGE_N], 1ull);
          if (i < P.range_cap)
          {
            P.range[i].x = (float)L.px;
            P.range[i].Z = el.Z;
          }
        }
        if (tally_on<TR>(P, MTB_TALLY_PHONON))
          L.casEnuc += den; // recoil.E + Elbind — TrimPhononOut::dissipateRecoilEnergy
        if (E2 < L.Ef)
          state = MTB_INTERSTITIAL;
      }
      // TrimPhononOut::checkPKAState — trim.C:503-511
      if (tally_on<TR>(P, MTB_TALLY_PHONON) && state != MTB_MOVING)
        L.casEnuc += L.E;
    }

    if (EVENTS)
    {
      if (n_events < P.events_cap)
      {
        mtb_event & ev = P.events[(size_t)lane_global * P.events_cap + n_events];
        ev.pka_pos[0] = MTB_AHEAD_X; ev.pka_pos[1] = MTB_AHEAD_Y; ev.pka_pos[2] = MTB_AHEAD_Z;
        ev.pka_dir[0] = L.dx; ev.pka_dir[1] = L.dy; ev.pka_dir[2] = L.dz;
        ev.pka_E = L.E;
        ev.recoil_pos[0] = L.px; ev.recoil_pos[1] = L.py; ev.recoil_pos[2] = L.pz;
        float qs = 1.0f;
        if (above)
          qs = frsqrt(qx * qx + qy * qy + qz * qz);
        ev.recoil_dir[0] = qx * qs; ev.recoil_dir[1] = qy * qs; ev.recoil_dir[2] = qz * qs;
        ev.recoil_E = (double)Erec;
        ev.ls = (double)ls;
        ev.dee = dee;
        ev.den = den;
        // the hooks see the MaterialBase object of the LAYER (sample_layers.C:26-49), also where identical
        // materials of a stack were folded into one on the device
        int user_material = M.user_index;
        if (!TR::has(F_MONO) && P.geom_kind == MTB_GEOM_LAYERS && P.n_input_materials != P.n_materials)
        {
          int lo = 0, hi = P.n_layers - 1;
          while (lo < hi)
          {
            const int mid = (lo + hi) >> 1;
            if (L.px < S.layer_cum[mid])
              hi = mid;
            else
              lo = mid + 1;
          }
          user_material = lo < P.n_input_materials ? lo : P.n_input_materials - 1;
        }
        ev.material = user_material;
        ev.element = nn;
        ev.material_tag = mtag;
        ev.pka_state = state;
        ev.recoil_above_threshold = above ? 1 : 0;
        ev._pad = 0;
      }
      ++n_events;
      if (state != MTB_MOVING)
      {
        finish_ion<TR>(P, S, L, rows, state, MTB_AHEAD_X, MTB_AHEAD_Y, MTB_AHEAD_Z);
        active = false;
      }
      L.px = MTB_AHEAD_X;
      L.py = MTB_AHEAD_Y;
      L.pz = MTB_AHEAD_Z;
      break;
    }

    // ---------------- who flies next ----------------
    if (DEFER)
    {
      if (follow || state != MTB_MOVING)
      {
        // park the result; the hand-over phase at the loop head picks it up (L.p* stays the collision site)
        pend = 1u | (follow ? 2u : 0u) | ((uint32_t)state << 2);
        sqx = qx; sqy = qy; sqz = qz;
        sErec = Erec;
        smx = mfx; smy = mfy; smz = mfz;
        sw3 = w[3];
        srp = (uint32_t)(SPECIES_CLASS0 + el.tcls) | ((uint32_t)rec_gen << GEN_SHIFT);
        smtag = mtag;
        sdsafe = dsafe_here;
      }
      else
      {
        L.px = MTB_AHEAD_X;
        L.py = MTB_AHEAD_Y;
        L.pz = MTB_AHEAD_Z;
      }
    }
    else if (follow)
    {
      L.casIons++;
      const float qs = frsqrt(qx * qx + qy * qy + qz * qz);
      const uint64_t ruid = child_uid(L.uid, L.ic, w[3]);
      const uint32_t rpacked = (uint32_t)(SPECIES_CLASS0 + el.tcls) | ((uint32_t)rec_gen << GEN_SHIFT);
      const bool keep_projectile = (state == MTB_MOVING) && (E2 <= Erec);
      if (state == MTB_MOVING && !keep_projectile)
      {
        // both move on and the recoil has less energy: suspend the projectile (at its new position), fly
        // the recoil
        Lane T = L;
        T.px = MTB_AHEAD_X;
        T.py = MTB_AHEAD_Y;
        T.pz = MTB_AHEAD_Z;
        suspend_ion<TR>(P, S, sp, T, L.prim, Erec);
      }
      if (state != MTB_MOVING)
        finish_ion<TR>(P, S, L, rows, state, MTB_AHEAD_X, MTB_AHEAD_Y, MTB_AHEAD_Z);
      if (keep_projectile)
      {
        // suspend the recoil instead
        Lane R;
        R.px = L.px; R.py = L.py; R.pz = L.pz;
        R.E = (double)Erec;
        R.dx = qx * qs; R.dy = qy * qs; R.dz = qz * qs;
        R.ic = 0;
        R.uid = ruid;
        R.packed = rpacked;
        R.tag = mtag;
        R.dsafe = dsafe_here;
        if (tally_on<TR>(P, MTB_TALLY_IONLOG))
        {
          R.prim = L.prim;
          log_birth<TR>(P, R, el.Z);
        }
        suspend_ion<TR>(P, S, sp, R, L.prim, E2);
        L.px = MTB_AHEAD_X;
        L.py = MTB_AHEAD_Y;
        L.pz = MTB_AHEAD_Z;
      }
      else
      {
        // the recoil starts where L.p* still is
        L.E = (double)Erec;
        L.Ecur = Erec;
        L.dx = qx * qs; L.dy = qy * qs; L.dz = qz * qs;
        L.ic = 0;
        L.uid = ruid;
        L.packed = rpacked;
        L.tag = mtag;
        L.pcls = el.tcls;
        L.dsafe = dsafe_here;
        log_birth<TR>(P, L, el.Z);
      }
      if (TR::has(F_DIAG) && stack_depth(sp) > sp_max)
        sp_max = stack_depth(sp); // diagnostic high-water mark (generic kernels only)
    }
    else if (state != MTB_MOVING)
    {
      finish_ion<TR>(P, S, L, rows, state, MTB_AHEAD_X, MTB_AHEAD_Y, MTB_AHEAD_Z);
      active = false;
    }
    else
    {
      L.px = MTB_AHEAD_X;
      L.py = MTB_AHEAD_Y;
      L.pz = MTB_AHEAD_Z;
    }
#undef MTB_AHEAD_X
#undef MTB_AHEAD_Y
#undef MTB_AHEAD_Z
      } while (0);

    if (EVENTS && !active)
      done = true; // single-ion mode ends with the ion
  }

  if (!EVENTS)
    MTB_ATOMIC_MAX(&S.blk_u64[CNT_STACKMAX], (unsigned long long)sp_max);
  else if (P.event_counts)
  {
    if (lane_global < P.n_primaries)
      P.event_counts[lane_global] = (uint32_t)(n_events < 0xFFFFFFFFull ? n_events : 0xFFFFFFFFull);
  }
  else if (lane_global == 0)
    P.u64[CNT_EVENTS_N] = n_events;
}

} // namespace mtb
#endif
