// mtb_tables.h — host-side flattening of the MyTRIM plugin objects into device tables.
//
// Pure C++ (no CUDA calls): mtb_engine.cu uploads the vectors; tests/hostsim.cpp points the
// kernels' LaunchParams straight at them.  All derived constants are computed in double and
// rounded to float once.
#ifndef MTB_TABLES_H
#define MTB_TABLES_H

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "mtb_types.h"

namespace mtb
{

struct ZblRow
{
  double mm1, m1, mnat, rho, atrho, vfermi, heat, lfctr;
  double pcoef[8];
};

inline const ZblRow *
builtin_zbl()
{
  static const ZblRow rows[MTB_NZ] = {
#include "zbl_tables.inc"
  };
  return rows;
}

constexpr int kSmemHistMax = 2048; // depth bins mirrored per CTA in shared memory (2 x 8 KB)

// ZBL proton stopping in double — MaterialBase::rpstop, material.C:133-158
inline double
host_proton_stopping(const ZblRow & row, int z2, double e)
{
  const double pe = std::max(25.0, e);
  const double sl = row.pcoef[0] * std::pow(pe, row.pcoef[1]) + row.pcoef[2] * std::pow(pe, row.pcoef[3]);
  const double sh = row.pcoef[4] / std::pow(pe, row.pcoef[5]) * std::log(row.pcoef[6] / pe + row.pcoef[7] * pe);
  double sp = sl * sh / (sl + sh);
  if (e <= 25.0)
    sp *= std::pow(e / 25.0, z2 <= 6 ? 0.25 : 0.45);
  return sp;
}

// Constants of the velocity-proportional heavy-ion regime (material.C:213-273 with yr == its
// lower clamp): effective charge zeta, matching energy eee, exponent.
inline LowStop
host_low_velocity_stopping(const ZblRow * zbl, int z1, int z2)
{
  LowStop out = {0.f, 0.f, 0.5f, 0.f};
  if (z1 < 3)
    return out;
  const double fz1 = z1;
  const double vfermi = zbl[z2 - 1].vfermi, lfctr = zbl[z1 - 1].lfctr;
  const double cb = std::cbrt(fz1), cb2 = cb * cb;
  const double yr = std::max(0.13, 1.0 / cb2);
  const double yr03 = std::pow(yr, 0.3);
  const double a = -0.803 * yr03 + 1.3167 * yr03 * yr03 + 0.38157 * yr + 0.008983 * yr * yr;
  const double q = std::min(1.0, std::max(0.0, 1.0 - std::exp(-std::min(a, 50.0))));
  const double b = std::min(0.43, std::max(0.32, 0.12 + 0.025 * fz1)) / cb;
  const double l0 = (0.8 - q * std::min(1.2, 0.6 + fz1 / 30.0)) / cb;
  const double qa = std::max(0.0, 0.9 - 0.025 * fz1), z16 = 0.025 * std::min(16.0, fz1);
  double l1;
  if (q < 0.2)
    l1 = 0.0;
  else if (q < qa)
    l1 = b * (q - 0.2) / std::abs(qa - 0.2000001);
  else if (q < std::max(0.0, 1.0 - z16))
    l1 = b;
  else
    l1 = b * (1.0 - q) / z16;
  const double l = std::max(l1, l0 * lfctr);
  const double lx = 4.0 * l * vfermi / 1.919;
  const double zeta = q + (1.0 / (2.0 * vfermi * vfermi)) * (1.0 - q) * std::log(1.0 + lx * lx);
  const double vrmin = std::max(1.0, 0.13 * cb2);
  const double vmin = 0.5 * (vrmin + std::sqrt(std::max(0.0, vrmin * vrmin - 0.8 * vfermi * vfermi)));
  const double eee = 25.0 * vmin * vmin;
  const double power = (z2 == 6 || ((z2 == 14 || z2 == 32) && z1 <= 19)) ? 0.375 : 0.5;
  const double zf = zeta * fz1;
  // e at which vr(e) / Z1^(2/3) leaves the clamp: vr is monotonic in e, bisect vr(e) = vrmin
  auto vr_of = [&](double e) {
    const double v = std::sqrt(e / 25.0) / vfermi, v2 = v * v;
    return v >= 1.0 ? v * vfermi * (1.0 + 1.0 / (5.0 * v2)) : (3.0 * vfermi / 4.0) * (1.0 + (2.0 * v2 / 3.0) - v2 * v2 / 15.0);
  };
  double e_sw = 0.0;
  if (vr_of(0.0) <= vrmin)
  {
    double lo = 0.0, hi = 1.0e7;
    for (int it = 0; it < 200; ++it)
    {
      const double mid = 0.5 * (lo + hi);
      if (vr_of(mid) <= vrmin)
        lo = mid;
      else
        hi = mid;
    }
    e_sw = lo;
  }
  // stay a hair inside the regime so that float rounding of e never selects the shortcut beyond it,
  // and below 20 keV/amu where the Z1^3 factor of material.C:254-255 is 1 to better than 1e-10
  out.e_max = (float)(std::min(e_sw, 20.0) * (1.0 - 1e-6));
  out.coef = (float)(10.0 * host_proton_stopping(zbl[z2 - 1], z2, eee) * zf * zf / std::pow(eee, power));
  out.power = (float)power;
  return out;
}

struct HostConfig
{
  mtb_config cfg;
  ZblRow zbl[MTB_NZ];
  std::vector<mtb_material> materials;
  std::vector<mtb_element> elements;
  mtb_geometry geom;
  std::vector<double> layer_thickness, cluster_xyzr;
  // (Z, m) of primaries seen so far that are not target atoms: they get projectile classes too
  std::vector<std::pair<int, double>> primary_species;
  std::vector<int> first_input_material; // set by fold_identical_materials on build_host_tables' private copy

  HostConfig()
  {
    std::memset(&cfg, 0, sizeof(cfg));
    std::memcpy(zbl, builtin_zbl(), sizeof(zbl));
    std::memset(&geom, 0, sizeof(geom));
    geom.kind = MTB_GEOM_SOLID;
    for (int i = 0; i < 3; ++i)
    {
      geom.w[i] = 10000.0; // sample.h:37
      geom.bc[i] = MTB_BC_PBC;
    }
  }
};

struct HostTables
{
  std::vector<DevElement> elements;
  std::vector<DevMaterial> materials;
  std::vector<DevIonZ> ionz;
  std::vector<LowStop> lowstop; // [MTB_NZ + 1][n_zslots]
  std::vector<ProjClass> pclass;
  std::vector<PairM> pairm;
  std::vector<PairE> paire;
  std::vector<int32_t> tclass_elem;
  std::vector<double> layer_cum;
  std::vector<int32_t> layer_mat, cl_hash, cl_next;
  std::vector<float> cl_safe; // per hash cell: distance bound to the nearest cluster surface (periodic boxes)
  std::vector<int> mat_map; // device material of every input material (fold_identical_materials)
  std::vector<uint8_t> cl_dist;
};

inline void
default_config(mtb_config * cfg)
{
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->tmin = 0.2; // simconf.C:44-49
  cfg->tau = 0.0;
  cfg->cw = 0.001;
  cfg->length_scale = 1.0;
  cfg->potential = MTB_POT_UNIVERSAL;
  cfg->follow = MTB_FOLLOW_ALL;
  cfg->follow_max_gen = 1;
  cfg->vacancy_model = MTB_VAC_COUNT;
  cfg->vmap_z[0] = cfg->vmap_z[1] = cfg->vmap_z[2] = -1;
}

inline int
check_materials(int n_materials, const mtb_material * materials, int n_elements, const mtb_element * elements,
                std::string & err)
{
  if (!materials || !elements || n_materials < 1 || n_elements < 1)
  {
    err = "bad materials";
    return MTB_EINVAL;
  }
  for (int i = 0; i < n_materials; ++i)
  {
    const mtb_material & m = materials[i];
    if (m.n_elements < 1 || m.first_element < 0 || m.first_element + m.n_elements > n_elements || !(m.rho > 0.0))
    {
      err = "bad material " + std::to_string(i);
      return MTB_EINVAL;
    }
  }
  for (int i = 0; i < n_elements; ++i)
    if (elements[i].Z < 1 || elements[i].Z > MTB_NZ || !(elements[i].m > 0.0))
    {
      err = "bad element " + std::to_string(i);
      return MTB_EINVAL;
    }
  return MTB_OK;
}

inline int
check_geometry(const mtb_geometry * g, std::string & err)
{
  if (!g || g->kind < MTB_GEOM_SOLID || g->kind > MTB_GEOM_CLUSTERS)
  {
    err = "unknown geometry kind";
    return MTB_EINVAL;
  }
  if (g->kind == MTB_GEOM_LAYERS && (g->n_layers < 1 || !g->layer_thickness))
  {
    err = "layers geometry without layers";
    return MTB_EINVAL;
  }
  if (g->kind == MTB_GEOM_CLUSTERS &&
      (g->kn[0] < 1 || g->kn[1] < 1 || g->kn[2] < 1 || g->n_clusters < 0 || (g->n_clusters && !g->cluster_xyzr)))
  {
    err = "bad clusters geometry";
    return MTB_EINVAL;
  }
  return MTB_OK;
}

inline double
c_tmin(const HostConfig & H)
{
  return H.cfg.tmin;
}

inline bool
species_known(const HostConfig & H, int Z, double m)
{
  const std::pair<int, double> key(Z, m);
  if (std::find(H.primary_species.begin(), H.primary_species.end(), key) != H.primary_species.end())
    return true;
  for (const auto & e : H.elements)
    if (e.Z == Z && e.m == m)
      return true;
  return false;
}

// Registers the distinct (Z, m) among the first `n` primaries given as projectile classes (at most
// `cap` beyond the target atoms).  This is only an optimisation for beams: a primary whose species
// has no class builds its own rows on the device when it is fetched.  Returns true if the tables
// have to be rebuilt.
inline bool
register_primary_species(HostConfig & H, uint64_t n, const mtb_ion * ions, size_t cap = 16)
{
  bool changed = false;
  int lastZ = -1;
  double lastM = -1.0;
  for (uint64_t i = 0; i < n; ++i)
  {
    if (ions[i].Z == lastZ && ions[i].m == lastM)
      continue;
    lastZ = ions[i].Z;
    lastM = ions[i].m;
    const std::pair<int, double> key(lastZ, lastM);
    bool known = std::find(H.primary_species.begin(), H.primary_species.end(), key) != H.primary_species.end();
    for (size_t e = 0; !known && e < H.elements.size(); ++e)
      known = H.elements[e].Z == key.first && H.elements[e].m == key.second;
    if (known)
      continue;
    if (H.primary_species.size() >= cap)
      return changed; // further species build their rows on the device
    if (key.first < 1 || key.first > MTB_NZ)
      continue;
    H.primary_species.push_back(key);
    changed = true;
  }
  return changed;
}

// Fills T and every non-pointer field of P.  Pointer fields of P are left for the caller.
// Layer stacks repeat materials (inputs/samplelayers_zro2_multilayer.in: 50 layers of one ZrO2): identical materials
// (density, tag, element list) are folded into one, so that such a stack has ONE material on the device — no layer
// search at all (LaunchParams::one_material), class tables of one material instead of fifty.  Only for solid/layered
// samples: in the wire and clusters geometries the material index itself carries meaning.  map[i] = device material
// of input material i.
inline std::vector<int>
fold_identical_materials(HostConfig & H)
{
  const int nm = (int)H.materials.size();
  std::vector<int> map(nm);
  for (int i = 0; i < nm; ++i)
    map[i] = i;
  if (H.geom.kind != MTB_GEOM_SOLID && H.geom.kind != MTB_GEOM_LAYERS)
    return map;
  auto same = [&H](const mtb_material & a, const mtb_material & b) {
    if (a.rho != b.rho || a.tag != b.tag || a.n_elements != b.n_elements)
      return false;
    for (int j = 0; j < a.n_elements; ++j)
    {
      const mtb_element & x = H.elements[a.first_element + j];
      const mtb_element & y = H.elements[b.first_element + j];
      if (x.Z != y.Z || x.m != y.m || x.t != y.t || x.Edisp != y.Edisp || x.Elbind != y.Elbind)
        return false;
    }
    return true;
  };
  std::vector<mtb_material> mats;
  std::vector<mtb_element> els;
  std::vector<int> first_input; // input index of the first occurrence
  for (int i = 0; i < nm; ++i)
  {
    int found = -1;
    for (size_t k = 0; k < first_input.size() && found < 0; ++k)
      if (same(H.materials[first_input[k]], H.materials[i]))
        found = (int)k;
    if (found < 0)
    {
      found = (int)mats.size();
      mtb_material m = H.materials[i];
      m.first_element = (int)els.size();
      els.insert(els.end(), H.elements.begin() + H.materials[i].first_element,
                 H.elements.begin() + H.materials[i].first_element + H.materials[i].n_elements);
      mats.push_back(m);
      first_input.push_back(i);
    }
    map[i] = found;
  }
  if ((int)mats.size() < nm)
  {
    H.materials.swap(mats);
    H.elements.swap(els);
  }
  else
    first_input.clear();
  H.first_input_material = first_input;
  return map;
}

inline int
build_host_tables(const HostConfig & H_in, HostTables & T, LaunchParams & P, std::string & err)
{
  std::memset(&P, 0, sizeof(P));
  if (H_in.materials.empty())
  {
    err = "mtb_set_materials has not been called";
    return MTB_EINVAL;
  }
  HostConfig H = H_in;
  T.mat_map = fold_identical_materials(H);
  const mtb_config & c = H.cfg;
  P.tmin = (float)c.tmin;
  P.tau = (float)c.tau;
  P.cw = (float)c.cw;
  P.inv_scale = (float)(1.0 / c.length_scale);
  P.potential = c.potential;
  P.follow = c.follow;
  P.follow_max_gen = c.follow_max_gen;
  P.vacancy_model = c.vacancy_model;
  P.tally_mask = c.tally_mask;
  for (int i = 0; i < 3; ++i)
    P.vmap_z[i] = c.vmap_z[i];
  P.ionlog_z = c.ionlog_z;

  // materials: MaterialBase::prepare — material.C:36-74
  T.elements.assign(H.elements.size(), DevElement());
  T.materials.assign(H.materials.size(), DevMaterial());
  for (size_t i = 0; i < H.materials.size(); ++i)
  {
    const mtb_material & m = H.materials[i];
    double tt = 0.0;
    for (int j = 0; j < m.n_elements; ++j)
      tt += std::max(0.0, H.elements[m.first_element + j].t);
    if (!(tt > 0.0))
    {
      err = "material with zero stoichiometry";
      return MTB_EINVAL;
    }
    double am = 0.0, az = 0.0;
    for (int j = 0; j < m.n_elements; ++j)
    {
      const mtb_element & e = H.elements[m.first_element + j];
      const double t = std::max(0.0, e.t) / tt;
      am += e.m * t;
      az += (double)e.Z * t;
      const ZblRow & row = H.zbl[e.Z - 1];
      DevElement & d = T.elements[m.first_element + j];
      std::memset(&d, 0, sizeof(d));
      d.m = (float)e.m;
      d.t = (float)t;
      d.Edisp = (float)e.Edisp;
      d.Elbind = (float)e.Elbind;
      d.z023 = (float)std::pow((double)e.Z, 0.23);
      d.fz = (float)e.Z;
      d.vfermi = (float)row.vfermi;
      d.vf2inv = (float)(1.0 / (2.0 * row.vfermi * row.vfermi));
      for (int k = 0; k < 8; ++k)
        d.pc[k] = (float)row.pcoef[k];
      d.Z = e.Z;
      d.velpwr = e.Z <= 6 ? 0.25f : 0.45f;
      {
        // rpstop at pe = 25 keV/amu (material.C:137-150): below that energy the reference evaluates the same
        // expression at 25 and scales it with (e / 25)^velpwr, so the value is a constant of the element
        const double * pc = row.pcoef;
        const double sl = pc[0] * std::pow(25.0, pc[1]) + pc[2] * std::pow(25.0, pc[3]);
        const double sh = pc[4] / std::pow(25.0, pc[5]) * std::log(pc[6] / 25.0 + pc[7] * 25.0);
        d.sp25 = (float)(sl * sh / (sl + sh));
      }
    }
    DevMaterial & d = T.materials[i];
    d.arho = (float)(m.rho * 0.6022 / am);
    d.am = (float)am;
    d.az = (float)az;
    d.az023 = (float)std::pow(az, 0.23);
    d.n_elem = m.n_elements;
    d.first_elem = m.first_element;
    d.tag = m.tag;
    d.user_index = H.first_input_material.empty() ? (int32_t)i : (int32_t)H.first_input_material[i];
  }
  T.ionz.assign(MTB_NZ + 1, DevIonZ());
  std::memset(T.ionz.data(), 0, sizeof(DevIonZ) * T.ionz.size());
  for (int z = 1; z <= MTB_NZ; ++z)
  {
    T.ionz[z].z023 = (float)std::pow((double)z, 0.23);
    T.ionz[z].cbrt = (float)std::cbrt((double)z);
    T.ionz[z].lfctr = (float)H.zbl[z - 1].lfctr;
    T.ionz[z].mm1 = (float)H.zbl[z - 1].mm1;
  }
  // low-velocity stopping table over the distinct target Z of this configuration
  std::vector<int> zlist;
  for (auto & d : T.elements)
  {
    auto it = std::find(zlist.begin(), zlist.end(), d.Z);
    if (it == zlist.end())
    {
      zlist.push_back(d.Z);
      it = zlist.end() - 1;
    }
    d.zslot = (int32_t)(it - zlist.begin());
  }
  P.n_zslots = (int32_t)zlist.size();
  T.lowstop.assign((size_t)(MTB_NZ + 1) * zlist.size(), LowStop{0.f, 0.f, 0.5f, 0.f});
  for (int z1 = 1; z1 <= MTB_NZ; ++z1)
    for (size_t k = 0; k < zlist.size(); ++k)
      T.lowstop[(size_t)z1 * zlist.size() + k] = host_low_velocity_stopping(H.zbl, z1, zlist[k]);
  // projectile / target classes and their pair tables — MaterialBase::average, material.C:77-110
  std::vector<std::pair<int, double>> classes;
  for (size_t i = 0; i < H.elements.size(); ++i)
  {
    const std::pair<int, double> key(H.elements[i].Z, H.elements[i].m);
    auto it = std::find(classes.begin(), classes.end(), key);
    if (it == classes.end())
    {
      classes.push_back(key);
      it = classes.end() - 1;
    }
    T.elements[i].tcls = (int32_t)(it - classes.begin());
  }
  const size_t nt = classes.size();
  T.tclass_elem.assign(nt, 0);
  for (size_t i = H.elements.size(); i-- > 0;)
    T.tclass_elem[T.elements[i].tcls] = (int32_t)i;
  for (const auto & sp : H.primary_species)
    if (std::find(classes.begin(), classes.end(), sp) == classes.end())
      classes.push_back(sp);
  const size_t np = classes.size(), nm = H.materials.size();
  T.pclass.assign(np, ProjClass());
  T.pairm.assign(np * nm, PairM());
  T.paire.assign(np * nt, PairE());
  for (size_t pc = 0; pc < np; ++pc)
  {
    const int z1 = classes[pc].first;
    const double m1 = classes[pc].second == 0.0 ? H.zbl[z1 - 1].mm1 : classes[pc].second;
    const double z1p = std::pow((double)z1, 0.23);
    ProjClass & c = T.pclass[pc];
    c.m2 = (float)(2.0 * m1);
    c.inv_km = (float)(0.001 / m1);
    c.m = (float)m1;
    c.fz = (float)z1;
    c.z023 = (float)z1p;
    c.cbrt = (float)std::cbrt((double)z1);
    c.lfctr = (float)H.zbl[z1 - 1].lfctr;
    c.Z = z1;
    for (size_t mi = 0; mi < nm; ++mi)
    {
      const mtb_material & m = H.materials[mi];
      double tt = 0.0, am = 0.0, az = 0.0;
      for (int j = 0; j < m.n_elements; ++j)
        tt += std::max(0.0, H.elements[m.first_element + j].t);
      for (int j = 0; j < m.n_elements; ++j)
      {
        const mtb_element & e = H.elements[m.first_element + j];
        am += e.m * std::max(0.0, e.t) / tt;
        az += (double)e.Z * std::max(0.0, e.t) / tt;
      }
      const double arho = m.rho * 0.6022 / am;
      const double mu = m1 / am;
      const double a = .5292 * .8853 / (z1p + std::pow(az, 0.23));
      const double f = a * am / (az * (double)z1 * 14.4 * (m1 + am));
      const double epsdg = c_tmin(H) * f * (1.0 + mu) * (1.0 + mu) / (4.0 * mu);
      PairM & pm = T.pairm[pc * nm + mi];
      pm.a = (float)a;
      pm.K = (float)std::sqrt(f * epsdg);
      pm.C2 = (float)(1.0 / (3.14159265358979323846 * arho * a * a));
      pm.sk = (float)std::sqrt(0.001 / m1);
    }
    for (size_t tc = 0; tc < nt; ++tc)
    {
      const int z2 = classes[tc].first;
      const double m2 = classes[tc].second;
      const double my = m1 / m2;
      const double ai = .5292 * .8853 / (z1p + std::pow((double)z2, 0.23));
      PairE & pe = T.paire[pc * nt + tc];
      pe.my = (float)my;
      pe.ec = (float)(4.0 * my / ((1.0 + my) * (1.0 + my)));
      pe.inv_ai = (float)(1.0 / ai);
      pe.sfi = (float)std::sqrt(ai * m2 / ((double)z1 * (double)z2 * 14.4 * (m1 + m2)));
    }
  }
  P.n_pclass = (int32_t)np;
  P.n_tclass = (int32_t)nt;
  P.n_elements = (int32_t)T.elements.size();
  P.n_materials = (int32_t)T.materials.size();
  P.n_input_materials = (int32_t)T.mat_map.size();
  if (P.n_pclass + SPECIES_CLASS0 > SPECIES_MASK)
  {
    err = "too many elements";
    return MTB_EINVAL;
  }

  // geometry
  const mtb_geometry & g = H.geom;
  P.geom_kind = g.kind;
  for (int i = 0; i < 3; ++i)
  {
    P.bc[i] = g.bc[i];
    P.w[i] = g.w[i];
  }
  double extent = g.w[0];
  T.layer_cum.clear();
  T.layer_mat.clear();
  T.cl_hash.clear();
  T.cl_next.clear();
  if (g.kind == MTB_GEOM_LAYERS)
  {
    const int nl = (int)H.layer_thickness.size();
    double d = 0.0;
    for (int i = 0; i < nl; ++i)
    {
      d += H.layer_thickness[i]; // same running sum as sample_layers.C:32-37
      T.layer_cum.push_back(d);
      // a position beyond the last interface belongs to the last material (sample_layers.C:40-41)
      T.layer_mat.push_back(T.mat_map[std::min(i, (int)T.mat_map.size() - 1)]);
    }
    P.n_layers = nl;
    extent = std::max(extent, d);
  }
  else if (g.kind == MTB_GEOM_BURIED_WIRE || g.kind == MTB_GEOM_CLUSTERS)
  {
    if (P.n_materials < 2)
    {
      err = "this geometry needs two materials";
      return MTB_EINVAL;
    }
  }
  if (g.kind == MTB_GEOM_CLUSTERS)
  {
    // sampleClusters::initSpatialhash + addCluster — sample_clusters.C:135-213
    const size_t ncell = (size_t)g.kn[0] * g.kn[1] * g.kn[2];
    const size_t ncl = H.cluster_xyzr.size() / 4;
    T.cl_hash.assign(ncell, -1);
    T.cl_next.assign(std::max<size_t>(ncl, 1), -1);
    double cmr = 0.0;
    for (size_t cidx = 0; cidx < ncl; ++cidx)
    {
      int k[3];
      for (int i = 0; i < 3; ++i)
      {
        k[i] = (int)std::floor((H.cluster_xyzr[4 * cidx + i] * g.kn[i]) / g.w[i]) % g.kn[i];
        if (k[i] < 0)
          k[i] += g.kn[i];
      }
      const size_t cell = (size_t)k[0] + (size_t)g.kn[0] * ((size_t)k[1] + (size_t)g.kn[1] * (size_t)k[2]);
      if (T.cl_hash[cell] < 0)
        T.cl_hash[cell] = (int32_t)cidx;
      else
      {
        int l = T.cl_hash[cell];
        while (T.cl_next[l] >= 0)
          l = T.cl_next[l];
        T.cl_next[l] = (int32_t)cidx;
      }
      cmr = std::max(cmr, H.cluster_xyzr[4 * cidx + 3]);
    }
    for (int i = 0; i < 3; ++i)
    {
      P.kn[i] = g.kn[i];
      P.kd[i] = g.w[i] / (double)g.kn[i];
      P.kn_w[i] = (double)g.kn[i] / g.w[i];
      P.inv_kn[i] = 1.0 / (double)g.kn[i];
      P.cl_ks[i] = (int)(cmr / P.kd[i]) + 1;
    }
    // Distance map in front of the scan.  lookupCluster scans the (2 ks + 1)^3 cells around the cell of
    // the position (sample_clusters.C:83-131).  Bubbles are sparse (7e-7 per A^3 in the uo2 case: 4 in
    // 59319 cells), so almost every scan visits only empty cells.  cl_dist[c] = 0 where a scan centred on
    // cell c can see a non-empty cell; elsewhere the chessboard distance (26-neighbourhood BFS, periodic
    // where the box is) to the nearest such cell.  The device scans only in cells with distance 0 and,
    // in a fully periodic box, lets an ion travel (distance - 1) cell edges before it looks again.
    T.cl_dist.assign(ncell, 255);
    std::vector<uint32_t> frontier, next;
    for (size_t cell = 0; cell < ncell; ++cell)
    {
      if (T.cl_hash[cell] < 0)
        continue;
      const int c[3] = {(int)(cell % g.kn[0]), (int)((cell / g.kn[0]) % g.kn[1]), (int)(cell / ((size_t)g.kn[0] * g.kn[1]))};
      std::vector<int> centres[3];
      for (int i = 0; i < 3; ++i)
        for (int d = -P.cl_ks[i]; d <= P.cl_ks[i]; ++d)
        {
          int k = c[i] - d; // a scan centred on k reaches c = k + d
          if (g.bc[i] == MTB_BC_PBC)
            k = ((k % g.kn[i]) + g.kn[i]) % g.kn[i];
          else if (k < 0 || k >= g.kn[i])
            continue;
          centres[i].push_back(k);
        }
      for (int k0 : centres[0])
        for (int k1 : centres[1])
          for (int k2 : centres[2])
          {
            const size_t cc = (size_t)k0 + (size_t)g.kn[0] * ((size_t)k1 + (size_t)g.kn[1] * (size_t)k2);
            if (T.cl_dist[cc])
            {
              T.cl_dist[cc] = 0;
              frontier.push_back((uint32_t)cc);
            }
          }
    }
    for (int level = 1; level < 255 && !frontier.empty(); ++level)
    {
      next.clear();
      for (uint32_t cell : frontier)
      {
        const int c[3] = {(int)(cell % g.kn[0]), (int)((cell / g.kn[0]) % g.kn[1]), (int)(cell / ((size_t)g.kn[0] * g.kn[1]))};
        for (int d0 = -1; d0 <= 1; ++d0)
          for (int d1 = -1; d1 <= 1; ++d1)
            for (int d2 = -1; d2 <= 1; ++d2)
            {
              int k[3] = {c[0] + d0, c[1] + d1, c[2] + d2};
              bool ok = true;
              for (int i = 0; i < 3; ++i)
              {
                if (g.bc[i] == MTB_BC_PBC)
                  k[i] = (k[i] + g.kn[i]) % g.kn[i];
                else if (k[i] < 0 || k[i] >= g.kn[i])
                  ok = false;
              }
              if (!ok)
                continue;
              const size_t cc = (size_t)k[0] + (size_t)g.kn[0] * ((size_t)k[1] + (size_t)g.kn[1] * (size_t)k[2]);
              if (T.cl_dist[cc] > level)
              {
                T.cl_dist[cc] = (uint8_t)level;
                next.push_back((uint32_t)cc);
              }
            }
      }
      frontier.swap(next);
    }
    // A path of length s changes the cell index by at most floor(s / kd) + 1 per axis, so an ion that
    // starts in a cell with distance n and travels less than (n - 1) kd_min ends in a cell with distance
    // >= 1: the lookup it skips would have returned the matrix.  Leaving a non-periodic box must be
    // seen by the lookup, so skipping is for fully periodic boxes only.
    const bool periodic = g.bc[0] == MTB_BC_PBC && g.bc[1] == MTB_BC_PBC && g.bc[2] == MTB_BC_PBC;
    P.cl_safe_unit = periodic ? (float)(0.999 * std::min(P.kd[0], std::min(P.kd[1], P.kd[2]))) : 0.f;
    P.cl_inv_safe_unit = P.cl_safe_unit > 0.f ? 1.0f / P.cl_safe_unit : 0.f;

    // Second map, in Angstrom: a lower bound of the distance from ANY point of a cell to the nearest cluster SURFACE.
    // The cell map above is blind within one scan range of a bubble (distance 0: scan at every step; distance 1: look
    // up at every step) although most of those cells are several Angstrom away from the sphere itself — and that is
    // where the recoils of a fission track near a bubble spend their lives.  An ion in such a cell may fly that far
    // before a look-up can find it inside a cluster, and a cell with a positive bound needs no scan at all.
    // Clusters whose window (scan range + 3 cells) does not reach a cell are more than 2 cell edges away from it, so
    // the bound is capped there.  Exact like the first map: only look-ups that return the matrix are skipped.
    T.cl_safe.clear();
    if (periodic && ncl > 0)
    {
      const double kd_min = std::min(P.kd[0], std::min(P.kd[1], P.kd[2]));
      T.cl_safe.assign(ncell, (float)(2.0 * kd_min));
      for (size_t l = 0; l < ncl; ++l)
      {
        double ctr[3];
        int cc[3], half[3];
        for (int i = 0; i < 3; ++i)
        {
          ctr[i] = H.cluster_xyzr[4 * l + i] - std::floor(H.cluster_xyzr[4 * l + i] / g.w[i]) * g.w[i];
          cc[i] = std::min(g.kn[i] - 1, (int)(ctr[i] / P.kd[i]));
          half[i] = std::min(P.cl_ks[i] + 3, g.kn[i] / 2);
        }
        const double rad = H.cluster_xyzr[4 * l + 3];
        for (int d0 = -half[0]; d0 <= half[0]; ++d0)
          for (int d1 = -half[1]; d1 <= half[1]; ++d1)
            for (int d2 = -half[2]; d2 <= half[2]; ++d2)
            {
              const int d[3] = {d0, d1, d2};
              int k[3];
              double dist2 = 0.0;
              for (int i = 0; i < 3; ++i)
              {
                k[i] = ((cc[i] + d[i]) % g.kn[i] + g.kn[i]) % g.kn[i];
                const double a = k[i] * P.kd[i], b = (k[i] + 1) * P.kd[i];
                double best = 1e300;
                for (int m = -1; m <= 1; ++m)
                {
                  const double x = ctr[i] + m * g.w[i];
                  best = std::min(best, std::max(0.0, std::max(a - x, x - b)));
                }
                dist2 += best * best;
              }
              const size_t cell = (size_t)k[0] + (size_t)g.kn[0] * ((size_t)k[1] + (size_t)g.kn[1] * (size_t)k[2]);
              const float bound = (float)std::max(0.0, 0.999 * (std::sqrt(dist2) - rad) - 1e-4);
              if (bound < T.cl_safe[cell])
                T.cl_safe[cell] = bound;
            }
      }
    }
  }

  // tally sizes
  int bins = c.hist_bins;
  if (bins <= 0)
    bins = (int)std::min<double>(std::max(16384.0, 4.0 * std::ceil(extent) + 1.0), 1 << 20);
  P.hist_bins = bins;
  P.evac_rows = c.evac_rows > 0 ? c.evac_rows : 32;
  int smem_hist_max = kSmemHistMax;
  if (const char * env = std::getenv("MYTRIM_B200_SMEM_HIST")) // tuning knob: depth bins mirrored in shared memory
    smem_hist_max = std::max(0, std::min(std::atoi(env), 8192));
  P.smem_hist_bins = (c.tally_mask & MTB_TALLY_VAC_DEPTH) ? std::min(bins, smem_hist_max) : 0;
  P.one_material = (P.n_materials == 1 && (P.geom_kind == MTB_GEOM_SOLID || P.geom_kind == MTB_GEOM_LAYERS)) ? 1 : 0;
  P.mono = (P.one_material && P.n_elements == 1) ? 1 : 0;
  P.ionlog_cap = (c.tally_mask & MTB_TALLY_IONLOG) ? (c.ionlog_capacity ? c.ionlog_capacity : (1ull << 20)) : 0;
  P.range_cap = (c.tally_mask & MTB_TALLY_RANGE) ? (c.range_capacity ? c.range_capacity : (1ull << 22)) : 0;
  return MTB_OK;
}

} // namespace mtb
#endif
