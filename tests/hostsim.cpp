// hostsim.cpp — TEST INFRASTRUCTURE: runs the device lane loop (mtb_transport.cuh, the very
// source the CUDA kernel is compiled from) single-threaded on the host.
//
// Purpose: there is no GPU in the build container, so control flow, stack handling, tallies and
// the FP32 reformulations are debugged here against the oracle before spending GPU time.  The
// MUFU approximations are replaced by libm float functions (mtb_math.cuh), so results differ
// from the GPU in the last bits but follow the same algorithm.  This is NOT part of the product
// and not a fallback: libmytrim_b200.so fails loudly without a CUDA device.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../mytrim_b200/csrc/mtb_tables.h"
#include "../mytrim_b200/csrc/mtb_transport.cuh"

using namespace mtb;

struct hs_engine
{
  HostConfig host;
  HostTables T;
  LaunchParams P;
  bool dirty = true;
  std::vector<unsigned long long> u64;
  double f64[2] = {0, 0};
  std::vector<unsigned int> hist;
  unsigned long long blk_u64[CNT_COUNT];
  double blk_f64[2];
  std::vector<StackEntry> stacks;
  std::vector<mtb_ion_log> ionlog;
  std::vector<RangeEntry> range;
  std::vector<float4_t> custom_rows;
  std::string err;
  bool force_generic = false;
};

static int
hs_prepare(hs_engine * e)
{
  if (!e->dirty)
    return MTB_OK;
  if (int rc = build_host_tables(e->host, e->T, e->P, e->err))
    return rc;
  LaunchParams & P = e->P;
  if (std::getenv("MYTRIM_B200_NO_MONO")) // same knob as mtb_engine.cu::build_tables
    P.mono = 0;
  P.elements = e->T.elements.data();
  P.materials = e->T.materials.data();
  P.ionz = e->T.ionz.data();
  P.lowstop = e->T.lowstop.data();
  P.pclass = e->T.pclass.data();
  P.pairm = e->T.pairm.data();
  P.paire = e->T.paire.data();
  P.layer_cum = e->T.layer_cum.data();
  P.layer_mat = e->T.layer_mat.data();
  P.cl_hash = e->T.cl_hash.data();
  P.cl_next = e->T.cl_next.data();
  P.cl_dist = e->T.cl_dist.data();
  P.cl_safe = e->T.cl_safe.empty() ? nullptr : e->T.cl_safe.data();
  P.cl_xyzr = e->host.cluster_xyzr.data();
  e->u64.assign(u64_block_size(P), 0ull);
  e->f64[0] = e->f64[1] = 0.0;
  e->hist.assign(2 * (size_t)P.smem_hist_bins + 1, 0u);
  e->stacks.assign(MTB_STACK_DEPTH, StackEntry());
  e->ionlog.assign(P.ionlog_cap, mtb_ion_log());
  e->range.assign(P.range_cap, RangeEntry());
  P.u64 = e->u64.data();
  P.f64 = e->f64;
  P.ionlog = e->ionlog.data();
  P.range = e->range.data();
  P.stacks = e->stacks.data();
  P.tclass_elem = e->T.tclass_elem.data();
  for (int pc = P.n_tclass; pc < P.n_pclass; ++pc) // as mtb_engine.cu: primary_class_rows_kernel
    primary_class_rows(pc, P.n_materials, P.n_tclass, P.tmin, e->T.ionz.data(), e->T.materials.data(), e->T.elements.data(),
                       e->T.tclass_elem.data(), e->T.pclass.data(), e->T.pairm.data(), e->T.paire.data());
  e->custom_rows.assign((size_t)(2 + P.n_materials + P.n_tclass), float4_t());
  P.custom_rows = e->custom_rows.data();
  e->dirty = false;
  return MTB_OK;
}

static BlockCtx
hs_ctx(hs_engine * e)
{
  BlockCtx S;
  S.elements = e->P.elements;
  S.materials = e->P.materials;
  S.ionz = e->P.ionz;
  S.lowstop = e->P.lowstop;
  S.pclass = e->P.pclass;
  S.pairm = e->P.pairm;
  S.paire = e->P.paire;
  S.layer_cum = e->P.layer_cum;
  S.layer_mat = e->P.layer_mat;
  S.hist_vac = e->hist.data();
  S.hist_repl = e->hist.data() + e->P.smem_hist_bins;
  S.blk_u64 = e->blk_u64;
  S.blk_f64 = e->blk_f64;
  S.pool = nullptr;
  S.pool_ctl = nullptr;
  std::memset(e->blk_u64, 0, sizeof(e->blk_u64));
  e->blk_f64[0] = e->blk_f64[1] = 0.0;
  std::fill(e->hist.begin(), e->hist.end(), 0u);
  return S;
}

static void
hs_flush(hs_engine * e)
{
  const LaunchParams & P = e->P;
  for (int i = 0; i < P.smem_hist_bins; ++i)
  {
    P.u64[off_vac(P) + i] += e->hist[i];
    P.u64[off_repl(P) + i] += e->hist[P.smem_hist_bins + i];
  }
  for (int i = 0; i < CNT_STACKMAX; ++i)
    P.u64[i] += e->blk_u64[i];
  if (e->blk_u64[CNT_STACKMAX] > P.u64[CNT_STACKMAX])
    P.u64[CNT_STACKMAX] = e->blk_u64[CNT_STACKMAX];
  P.f64[0] += e->blk_f64[0];
  P.f64[1] += e->blk_f64[1];
}

extern "C" {

hs_engine *
hs_create(const mtb_config * cfg)
{
  hs_engine * e = new hs_engine();
  e->host.cfg = *cfg;
  return e;
}

void
hs_destroy(hs_engine * e)
{
  delete e;
}

const char *
hs_error(hs_engine * e)
{
  return e->err.c_str();
}

int
hs_set_materials(hs_engine * e, int nm, const mtb_material * m, int ne, const mtb_element * el)
{
  if (int rc = check_materials(nm, m, ne, el, e->err))
    return rc;
  e->host.materials.assign(m, m + nm);
  e->host.elements.assign(el, el + ne);
  e->dirty = true;
  return MTB_OK;
}

int
hs_set_geometry(hs_engine * e, const mtb_geometry * g)
{
  if (int rc = check_geometry(g, e->err))
    return rc;
  e->host.layer_thickness.clear();
  e->host.cluster_xyzr.clear();
  if (g->kind == MTB_GEOM_LAYERS)
    e->host.layer_thickness.assign(g->layer_thickness, g->layer_thickness + g->n_layers);
  if (g->kind == MTB_GEOM_CLUSTERS && g->n_clusters)
    e->host.cluster_xyzr.assign(g->cluster_xyzr, g->cluster_xyzr + 4 * (size_t)g->n_clusters);
  e->host.geom = *g;
  e->host.geom.layer_thickness = nullptr;
  e->host.geom.cluster_xyzr = nullptr;
  e->dirty = true;
  return MTB_OK;
}

int
hs_run(hs_engine * e, uint64_t n, const mtb_ion * primaries, uint64_t seed, uint64_t first_index, mtb_record * records)
{
  if (register_primary_species(e->host, std::min<uint64_t>(n, 64), primaries))
    e->dirty = true;
  if (int rc = hs_prepare(e))
    return rc;
  LaunchParams & P = e->P;
  P.primaries = primaries;
  P.n_primaries = n;
  P.first_index = first_index;
  P.key0 = (uint32_t)seed;
  P.key1 = (uint32_t)(seed >> 32);
  philox_round_keys(P.key0, P.key1, P.rk);
  P.records = records;
  P.u64[CNT_NEXT_PRIMARY] = 0;
  const BlockCtx S = hs_ctx(e);
  P.index_list = nullptr;
  P.deferred = nullptr;
  P.u64[CNT_DEFERRED] = 0;
  // same variant selection as mtb_engine.cu::launch_transport / run_deferred
  const Variant v = e->force_generic ? VARIANT_GENERIC : pick_variant(P, false);
  const Variant vc = e->force_generic ? VARIANT_GENERIC : pick_variant(P, true);
  auto run_variant = [&](Variant which) {
    if (which == VARIANT_FAST)
      lane_loop<TraitsFast>(P, S, 0);
    else if (which == VARIANT_MONO)
      lane_loop<TraitsMono>(P, S, 0);
    else if (which == VARIANT_CLUSTERS)
      lane_loop<TraitsClusters>(P, S, 0);
    else if (which == VARIANT_FAST_PHONON)
      lane_loop<TraitsFastPhonon>(P, S, 0);
    else if (which == VARIANT_MONO_EVAC)
      lane_loop<TraitsMonoEvac>(P, S, 0);
    else if (which == VARIANT_CLUSTERS_LOG)
      lane_loop<TraitsClustersLog>(P, S, 0);
    else if (which == VARIANT_LAYERS_PLAIN)
      lane_loop<TraitsLayersPlain>(P, S, 0);
    else if (which == VARIANT_LAYERS)
      lane_loop<TraitsLayers>(P, S, 0);
    else
      lane_loop<TraitsGeneric>(P, S, 0);
  };
  if (!(variant_features(v) & F_CUSTOM))
  {
    std::vector<uint32_t> deferred(n ? n : 1);
    P.deferred = deferred.data();
    run_variant(v);
    const unsigned long long nd = P.u64[CNT_DEFERRED];
    if (nd)
    {
      // same hand-over as mtb_engine.cu::run_deferred
      P.index_list = deferred.data();
      P.deferred = nullptr;
      P.n_primaries = nd;
      P.u64[CNT_NEXT_PRIMARY] = 0;
      run_variant(vc);
      P.index_list = nullptr;
    }
  }
  else
    run_variant(v);
  hs_flush(e);
  const unsigned long long bad = P.u64[CNT_ERROR];
  P.u64[CNT_ERROR] = 0;
  if (bad >> 32)
    e->err = "primaries were skipped";
  return (bad >> 32) ? MTB_EINVAL : (bad ? MTB_ESTACK : MTB_OK);
}

// what build_host_tables made of the sample: {device materials, one_material, mono, layers}
int
hs_sample_info(hs_engine * e, int32_t * out4)
{
  if (int rc = hs_prepare(e))
    return rc;
  out4[0] = e->P.n_materials;
  out4[1] = e->P.one_material;
  out4[2] = e->P.mono;
  out4[3] = e->P.n_layers;
  return MTB_OK;
}

// the variant pick_variant() selects for the configuration (same function as mtb_engine.cu: mtb_kernel_variant)
const char *
hs_variant(hs_engine * e)
{
  if (hs_prepare(e))
    return "";
  return variant_name(e->force_generic ? VARIANT_GENERIC : pick_variant(e->P, false));
}

void
hs_force_generic(hs_engine * e, int on)
{
  e->force_generic = on != 0;
}

int
hs_get_counters(hs_engine * e, mtb_counters * out)
{
  if (int rc = hs_prepare(e))
    return rc;
  const unsigned long long * c = e->P.u64;
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
  out->EelTotal = e->P.f64[0];
  out->EnucTotal = e->P.f64[1];
  return MTB_OK;
}

int
hs_get_vac_depth(hs_engine * e, uint64_t * vac, uint64_t * repl, size_t capacity, size_t * n_bins)
{
  if (int rc = hs_prepare(e))
    return rc;
  const LaunchParams & P = e->P;
  const size_t B = (size_t)P.hist_bins;
  size_t last = 0;
  for (size_t i = 0; i < B; ++i)
    if (P.u64[off_vac(P) + i] || P.u64[off_repl(P) + i])
      last = i + 1;
  if (n_bins)
    *n_bins = last;
  for (size_t i = 0; i < capacity; ++i)
  {
    vac[i] = i < B ? P.u64[off_vac(P) + i] : 0;
    repl[i] = i < B ? P.u64[off_repl(P) + i] : 0;
  }
  return MTB_OK;
}

int
hs_get_vac_energy(hs_engine * e, uint64_t * evac, size_t rows, size_t bins)
{
  if (int rc = hs_prepare(e))
    return rc;
  const LaunchParams & P = e->P;
  std::memset(evac, 0, rows * bins * sizeof(uint64_t));
  for (size_t r = 0; r < std::min<size_t>(P.evac_rows, rows); ++r)
    for (size_t x = 0; x < std::min<size_t>(P.hist_bins, bins); ++x)
      evac[r * bins + x] = P.u64[off_evac(P) + r * (size_t)P.hist_bins + x];
  return MTB_OK;
}

int
hs_get_vacmap(hs_engine * e, uint64_t * vmap)
{
  if (int rc = hs_prepare(e))
    return rc;
  std::memcpy(vmap, e->P.u64 + off_vmap(e->P), MTB_VMAP_NX * MTB_VMAP_NY * 3 * sizeof(uint64_t));
  return MTB_OK;
}

int
hs_get_range_list(hs_engine * e, float * x, int32_t * Z, size_t capacity, size_t * n)
{
  if (int rc = hs_prepare(e))
    return rc;
  const size_t cnt = (size_t)e->P.u64[CNT_RANGE_N];
  if (n)
    *n = cnt;
  for (size_t i = 0; i < cnt && i < capacity && i < e->range.size(); ++i)
  {
    x[i] = e->range[i].x;
    Z[i] = e->range[i].Z;
  }
  return MTB_OK;
}

int
hs_get_ion_log(hs_engine * e, mtb_ion_log * out, size_t capacity, size_t * n)
{
  if (int rc = hs_prepare(e))
    return rc;
  const size_t have = std::min<size_t>((size_t)e->P.u64[CNT_IONLOG_N], e->ionlog.size());
  std::unordered_map<uint64_t, size_t> birth;
  for (size_t i = 0; i < have; ++i)
    if (e->ionlog[i].state == -1)
      birth[e->ionlog[i].uid] = i;
  size_t m = 0;
  for (size_t i = 0; i < have; ++i)
  {
    if (e->ionlog[i].state == -1)
      continue;
    auto it = birth.find(e->ionlog[i].uid);
    if (it == birth.end())
      continue;
    if (m < capacity)
    {
      mtb_ion_log v = e->ionlog[i];
      std::memcpy(v.pos0, e->ionlog[it->second].pos0, sizeof(v.pos0));
      v.E0 = e->ionlog[it->second].E0;
      out[m] = v;
    }
    ++m;
  }
  if (n)
    *n = m;
  return MTB_OK;
}

int
hs_trim_one(hs_engine * e, mtb_ion * ion, uint64_t seed, uint64_t uid, int32_t * final_state, mtb_event * events,
            size_t capacity, size_t * n_events)
{
  if (register_primary_species(e->host, 1, ion))
    e->dirty = true;
  if (int rc = hs_prepare(e))
    return rc;
  LaunchParams P = e->P;
  P.primaries = nullptr;
  P.beam = *ion;
  P.n_primaries = 1;
  P.first_index = 0;
  P.single_uid = uid;
  P.key0 = (uint32_t)seed;
  P.key1 = (uint32_t)(seed >> 32);
  philox_round_keys(P.key0, P.key1, P.rk);
  P.records = nullptr;
  P.events = events;
  P.events_cap = capacity;
  P.tally_mask = 0;
  P.u64[CNT_EVENTS_N] = 0;
  const BlockCtx S = hs_ctx(e);
  lane_loop<TraitsEvents>(P, S, 0);
  const size_t cnt = (size_t)P.u64[CNT_EVENTS_N];
  if (n_events)
    *n_events = cnt;
  if (cnt > capacity)
    return MTB_ECAPACITY;
  if (cnt)
  {
    const mtb_event & last = events[cnt - 1];
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
hs_stopping(hs_engine * e, int material, size_t n, const int32_t * Z1, const double * m1, const double * E, double * out)
{
  if (int rc = hs_prepare(e))
    return rc;
  const BlockCtx S = hs_ctx(e);
  for (size_t i = 0; i < n; ++i)
  {
    const ProjClass pr = make_proj_class(S.ionz[Z1[i]], Z1[i], (float)m1[i]);
    out[i] = (double)material_stopping(pr, S.lowstop + Z1[i] * e->P.n_zslots, e->P.materials[e->T.mat_map[material]], e->P.elements,
                                       (float)E[i], fsqrt((float)E[i] * pr.inv_km));
  }
  return MTB_OK;
}

} // extern "C"
