/*
 * oracle.c — CPU restatement of MyTRIM's cascade-transport hot path (see oracle.h).
 * TEST INFRASTRUCTURE ONLY: never linked into, or called from, the product library.
 *
 * All arithmetic is IEEE double and deliberately keeps the reference's operand
 * order (no FMA contraction: build with -ffp-contract=off), so that in
 * ORC_RNG_MT19937 mode the results equal the compiled reference bit for bit.
 * Citations are relative to the reference tree.
 */
#include "oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------- */
/* tables                                                                     */
/* ------------------------------------------------------------------------- */

typedef struct {
  double mm1, m1, mnat, rho, atrho, vfermi, heat, lfctr;
  double pcoef[8];
} orc_zbl_row;

static const orc_zbl_row orc_builtin_zbl[MTB_NZ] = {
#include "../mytrim_b200/csrc/zbl_tables.inc"
};

/* ------------------------------------------------------------------------- */
/* random numbers                                                             */
/* ------------------------------------------------------------------------- */

/* std::mt19937 (Matsumoto & Nishimura MT19937, 32-bit), as seeded by
 * std::mt19937(seed) — simconf.C:40, 65-71. */
void
orc_mt_seed(orc_mt19937 * g, uint32_t seed)
{
  g->mt[0] = seed;
  for (int i = 1; i < 624; ++i)
    g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
  g->idx = 624;
}

uint32_t
orc_mt_next(orc_mt19937 * g)
{
  if (g->idx >= 624)
  {
    for (int k = 0; k < 624; ++k)
    {
      uint32_t y = (g->mt[k] & 0x80000000u) | (g->mt[(k + 1) % 624] & 0x7fffffffu);
      uint32_t v = g->mt[(k + 397) % 624] ^ (y >> 1);
      if (y & 1u)
        v ^= 0x9908b0dfu;
      g->mt[k] = v;
    }
    g->idx = 0;
  }
  uint32_t y = g->mt[g->idx++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

/* SimconfType::drand (simconf.h:52): uniform_real_distribution<double>(0,1) over mt19937 is
 * libstdc++'s generate_canonical<double,53>: two 32-bit draws, low word first, summed in
 * double, divided by 2^64, clamped below 1. */
double
orc_mt_drand(orc_mt19937 * g)
{
  double sum = (double)orc_mt_next(g);
  sum += (double)orc_mt_next(g) * 4294967296.0;
  double r = sum / 18446744073709551616.0;
  if (r >= 1.0)
    r = nextafter(1.0, 0.0);
  return r;
}

/* SimconfType::irand (simconf.h:53, simconf.C:42): uniform_int_distribution<unsigned>(0,65535)
 * over a 32-bit engine is Lemire's multiply-shift in libstdc++ >= 11: (x * 65536) >> 32; the
 * rejection threshold (2^32 mod 65536) is zero. */
uint32_t
orc_mt_irand(orc_mt19937 * g)
{
  return orc_mt_next(g) >> 16;
}

/* Philox4x32-10 (Salmon et al., Random123).  Constants as in curand_philox4x32_x.h:171-193. */
void
orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int round = 0; round < 10; ++round)
  {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Stream id of a recoil: spare word of the parent's Philox block of the collision that created it in
 * the high half, a mixed copy of (parent id, collision index) in the low half. */
uint64_t
orc_child_uid(uint64_t uid, uint32_t ic, uint32_t w3)
{
  const uint32_t lo = ((uint32_t)uid ^ (uint32_t)(uid >> 32)) * 0x9E3779B9u + ic * 0x85EBCA6Bu;
  return ((uint64_t)w3 << 32) | (uint64_t)lo;
}

/* 32 random bits -> uniform in (0,1): top 23 bits as the mantissa of a float in [1,2), minus 1, plus
 * half a grid step; computed in single precision exactly as on the device */
float
orc_u01(uint32_t x)
{
  union { uint32_t u; float f; } cvt;
  cvt.u = 0x3f800000u | (x >> 9);
  return (cvt.f - 1.0f) + 0x1p-24f;
}

/* ------------------------------------------------------------------------- */
/* engine state                                                               */
/* ------------------------------------------------------------------------- */

typedef struct {
  int Z;
  double m, t, Edisp, Elbind;
  /* ion dependent, set by average (material.C:99-108) */
  double my, ec, ai, fi;
} orc_element;

typedef struct {
  double rho, am, az, arho;
  int tag, n_elements;
  orc_element * el;
  /* ion dependent (material.C:80-93) */
  double mu, a, f, epsdg, pmax;
  int averaged;
} orc_material;

typedef struct {
  double pos[3], dir[3];
  double E, m, Ef;
  int Z, gen, tag, state, id;
  uint64_t uid, primary;
  uint32_t ic;
  int is_primary;
  double pos0[3], E0;
} orc_ion;

typedef struct {
  orc_ion * buf;
  size_t cap, head, count;
} orc_fifo;

typedef struct {
  uint64_t * v;
  size_t n;
} orc_hist;

struct orc_engine {
  mtb_config cfg;
  int rng_mode;
  orc_mt19937 mt;
  uint32_t key[2];

  orc_zbl_row zbl[MTB_NZ];

  int n_materials;
  orc_material * mat;

  mtb_geometry geom;
  double * layer_thickness;
  /* clusters (sample_clusters.h:29-53) */
  int * sh, *cl, cn, cnm, kn[3];
  double kd[3], cmr;
  double * c[4];
  int last_cluster;

  /* tallies */
  mtb_counters cnt;
  int vacancies_int; /* the reference's `int vacancies_created` arithmetic for NRT/KP */
  orc_hist vac, repl;
  orc_hist * evac;
  size_t evac_rows;
  uint64_t vmap[MTB_VMAP_NX * MTB_VMAP_NY * 3];
  double * range_x;
  int32_t * range_z;
  size_t range_n, range_cap;
  mtb_ion_log * ilog;
  size_t ilog_n, ilog_cap;
  int next_id;

  /* per-cascade */
  uint32_t cas_repl;

  /* event sink for orc_trim_one */
  mtb_event * ev;
  size_t ev_cap, ev_n;

  orc_fifo fifo;
};

static void
hist_add(orc_hist * h, size_t bin)
{
  if (bin >= h->n)
  {
    size_t n = bin + 1;
    h->v = (uint64_t *)realloc(h->v, n * sizeof(uint64_t));
    memset(h->v + h->n, 0, (n - h->n) * sizeof(uint64_t));
    h->n = n;
  }
  h->v[bin]++;
}

static void
fifo_push(orc_fifo * q, const orc_ion * ion)
{
  if (q->count == q->cap)
  {
    size_t ncap = q->cap ? 2 * q->cap : 256;
    orc_ion * nb = (orc_ion *)malloc(ncap * sizeof(orc_ion));
    for (size_t i = 0; i < q->count; ++i)
      nb[i] = q->buf[(q->head + i) % q->cap];
    free(q->buf);
    q->buf = nb;
    q->cap = ncap;
    q->head = 0;
  }
  q->buf[(q->head + q->count) % q->cap] = *ion;
  q->count++;
}

static int
fifo_pop(orc_fifo * q, orc_ion * out)
{
  if (!q->count)
    return 0;
  *out = q->buf[q->head];
  q->head = (q->head + 1) % q->cap;
  q->count--;
  return 1;
}

orc_engine *
orc_create(const mtb_config * cfg, int rng_mode)
{
  orc_engine * e = (orc_engine *)calloc(1, sizeof(orc_engine));
  e->cfg = *cfg;
  e->rng_mode = rng_mode;
  memcpy(e->zbl, orc_builtin_zbl, sizeof(orc_builtin_zbl));
  e->last_cluster = -1;
  e->geom.kind = MTB_GEOM_SOLID;
  for (int i = 0; i < 3; ++i)
  {
    e->geom.w[i] = 10000.0; /* sample.h:37 defaults */
    e->geom.bc[i] = MTB_BC_PBC;
  }
  return e;
}

static void
free_materials(orc_engine * e)
{
  for (int i = 0; i < e->n_materials; ++i)
    free(e->mat[i].el);
  free(e->mat);
  e->mat = 0;
  e->n_materials = 0;
}

static void
free_geometry(orc_engine * e)
{
  free(e->layer_thickness);
  e->layer_thickness = 0;
  free(e->sh);
  free(e->cl);
  e->sh = e->cl = 0;
  for (int i = 0; i < 4; ++i)
  {
    free(e->c[i]);
    e->c[i] = 0;
  }
  e->cn = e->cnm = 0;
}

int
orc_reset_tallies(orc_engine * e)
{
  memset(&e->cnt, 0, sizeof(e->cnt));
  e->vacancies_int = 0;
  free(e->vac.v);
  free(e->repl.v);
  memset(&e->vac, 0, sizeof(e->vac));
  memset(&e->repl, 0, sizeof(e->repl));
  for (size_t i = 0; i < e->evac_rows; ++i)
    free(e->evac[i].v);
  free(e->evac);
  e->evac = 0;
  e->evac_rows = 0;
  memset(e->vmap, 0, sizeof(e->vmap));
  e->range_n = 0;
  e->ilog_n = 0;
  e->next_id = 0;
  return MTB_OK;
}

void
orc_destroy(orc_engine * e)
{
  if (!e)
    return;
  orc_reset_tallies(e);
  free(e->range_x);
  free(e->range_z);
  free(e->ilog);
  free_materials(e);
  free_geometry(e);
  free(e->fifo.buf);
  free(e);
}

int
orc_set_tables(orc_engine * e, const double * pcoef, const double * vfermi, const double * lfctr,
               const double * mm1)
{
  for (int z = 0; z < MTB_NZ; ++z)
  {
    if (pcoef)
      memcpy(e->zbl[z].pcoef, pcoef + 8 * z, 8 * sizeof(double));
    if (vfermi)
      e->zbl[z].vfermi = vfermi[z];
    if (lfctr)
      e->zbl[z].lfctr = lfctr[z];
    if (mm1)
      e->zbl[z].mm1 = mm1[z];
  }
  return MTB_OK;
}

/* MaterialBase::prepare — material.C:36-74 */
int
orc_set_materials(orc_engine * e, int n_materials, const mtb_material * materials, int n_elements,
                  const mtb_element * elements)
{
  free_materials(e);
  e->mat = (orc_material *)calloc((size_t)n_materials, sizeof(orc_material));
  e->n_materials = n_materials;
  for (int i = 0; i < n_materials; ++i)
  {
    orc_material * M = &e->mat[i];
    const mtb_material * in = &materials[i];
    if (in->first_element < 0 || in->first_element + in->n_elements > n_elements)
      return MTB_EINVAL;
    M->rho = in->rho;
    M->tag = in->tag;
    M->n_elements = in->n_elements;
    M->el = (orc_element *)calloc((size_t)in->n_elements, sizeof(orc_element));
    double tt = 0.0;
    for (int j = 0; j < in->n_elements; ++j)
    {
      const mtb_element * s = &elements[in->first_element + j];
      orc_element * d = &M->el[j];
      if (s->Z < 1 || s->Z > MTB_NZ)
        return MTB_EINVAL;
      d->Z = s->Z;
      d->m = s->m;
      d->t = s->t < 0.0 ? 0.0 : s->t;
      d->Edisp = s->Edisp;
      d->Elbind = s->Elbind;
      tt += d->t;
    }
    for (int j = 0; j < M->n_elements; ++j)
      M->el[j].t /= tt;
    M->am = 0.0;
    M->az = 0.0;
    for (int j = 0; j < M->n_elements; ++j)
    {
      M->am += M->el[j].m * M->el[j].t;
      M->az += (double)M->el[j].Z * M->el[j].t;
    }
    M->arho = M->rho * 0.6022 / M->am; /* atoms/Ang^3 */
  }
  return MTB_OK;
}

/* sampleClusters::addCluster — sample_clusters.C:175-213 */
static void
clusters_add(orc_engine * e, double x, double y, double z, double r)
{
  if (e->cn >= e->cnm)
  {
    int n = e->cnm + e->cnm / 10 + 10;
    e->cl = (int *)realloc(e->cl, sizeof(int) * (size_t)n);
    for (int i = 0; i < 4; ++i)
      e->c[i] = (double *)realloc(e->c[i], sizeof(double) * (size_t)n);
    for (int j = e->cnm; j < n; ++j)
      e->cl[j] = -1;
    e->cnm = n;
  }
  const int cn = e->cn;
  e->c[0][cn] = x;
  e->c[1][cn] = y;
  e->c[2][cn] = z;
  e->c[3][cn] = r;
  int k[3];
  for (int i = 0; i < 3; ++i)
  {
    k[i] = (int)floor((e->c[i][cn] * e->kn[i]) / e->geom.w[i]) % e->kn[i];
    if (k[i] < 0)
      k[i] += e->kn[i];
  }
  int l = k[0] + e->kn[0] * (k[1] + e->kn[1] * k[2]);
  if (e->sh[l] < 0)
    e->sh[l] = cn;
  else
  {
    l = e->sh[l];
    while (e->cl[l] >= 0)
      l = e->cl[l];
    e->cl[l] = cn;
  }
  e->cl[cn] = -1;
  if (r > e->cmr)
    e->cmr = r;
  e->cn++;
}

/* sampleClusters::initSpatialhash — sample_clusters.C:135-152 */
static void
clusters_init_hash(orc_engine * e, int x, int y, int z)
{
  e->kn[0] = x;
  e->kn[1] = y;
  e->kn[2] = z;
  size_t n = (size_t)x * y * z;
  e->sh = (int *)malloc(sizeof(int) * n);
  for (size_t i = 0; i < n; ++i)
    e->sh[i] = -1;
  for (int i = 0; i < 3; ++i)
    e->kd[i] = e->geom.w[i] / (double)e->kn[i];
  e->cmr = 0.0;
}

int
orc_set_geometry(orc_engine * e, const mtb_geometry * g)
{
  free_geometry(e);
  e->geom = *g;
  e->geom.layer_thickness = 0;
  e->geom.cluster_xyzr = 0;
  if (g->kind == MTB_GEOM_LAYERS)
  {
    if (g->n_layers < 1 || !g->layer_thickness)
      return MTB_EINVAL;
    e->layer_thickness = (double *)malloc(sizeof(double) * (size_t)g->n_layers);
    memcpy(e->layer_thickness, g->layer_thickness, sizeof(double) * (size_t)g->n_layers);
  }
  if (g->kind == MTB_GEOM_CLUSTERS)
  {
    if (g->kn[0] < 1 || g->kn[1] < 1 || g->kn[2] < 1)
      return MTB_EINVAL;
    clusters_init_hash(e, g->kn[0], g->kn[1], g->kn[2]);
    for (int i = 0; i < g->n_clusters; ++i)
      clusters_add(e, g->cluster_xyzr[4 * i], g->cluster_xyzr[4 * i + 1], g->cluster_xyzr[4 * i + 2],
                   g->cluster_xyzr[4 * i + 3]);
  }
  return MTB_OK;
}

/* ------------------------------------------------------------------------- */
/* material physics                                                           */
/* ------------------------------------------------------------------------- */

/* MaterialBase::average — material.C:77-110 */
static void
material_average(const orc_engine * e, orc_material * M, int Z1, double m1)
{
  M->mu = m1 / M->am;
  const double fZ = (double)Z1;
  const double fZ023 = pow(fZ, 0.23);
  M->a = .5292 * .8853 / (fZ023 + pow(M->az, 0.23));
  M->f = M->a * M->am / (M->az * fZ * 14.4 * (m1 + M->am));
  M->epsdg = e->cfg.tmin * M->f * ((1.0 + M->mu) * (1.0 + M->mu)) / (4.0 * M->mu);
  for (int i = 0; i < M->n_elements; ++i)
  {
    orc_element * el = &M->el[i];
    el->my = m1 / el->m;
    el->ec = 4.0 * el->my / ((1.0 + el->my) * (1.0 + el->my));
    el->ai = .5292 * .8853 / (fZ023 + pow((double)el->Z, 0.23));
    el->fi = el->ai * el->m / (fZ * (double)el->Z * 14.4 * (m1 + el->m));
  }
  M->averaged = 1;
}

/* MaterialBase::rpstop — material.C:133-158: ZBL proton stopping, e in keV/amu */
static double
proton_stopping(const orc_engine * e, int z2p, double en)
{
  const double * pc = e->zbl[z2p - 1].pcoef;
  const double pe0 = 25.0;
  const double pe = pe0 > en ? pe0 : en; /* std::max(pe0, e) */
  const double sl = (pc[0] * pow(pe, pc[1])) + (pc[2] * pow(pe, pc[3]));
  const double sh = pc[4] / pow(pe, pc[5]) * log(pc[6] / pe + pc[7] * pe);
  double sp = sl * sh / (sl + sh);
  if (en <= pe0)
  {
    const double velpwr = z2p <= 6 ? 0.25 : 0.45;
    sp *= pow(en / pe0, velpwr);
  }
  return sp;
}

static inline double
dmax(double a, double b)
{
  return a < b ? b : a; /* std::max(a,b): returns a unless a < b */
}
static inline double
dmin(double a, double b)
{
  return b < a ? b : a; /* std::min(a,b) */
}

/* MaterialBase::rstop — material.C:160-282: electronic stopping cross-section of one target
 * element for ion (z1, m1, E) */
static double
element_stopping(const orc_engine * e, int z1, double m1in, double E, int z2)
{
  const double fz1 = (double)z1, fz2 = (double)z2;
  const double lfctr = e->zbl[z1 - 1].lfctr;
  const double mm1 = e->zbl[z1 - 1].mm1;
  const double vfermi = e->zbl[z2 - 1].vfermi;
  const double m1 = m1in == 0.0 ? mm1 : m1in;
  const double ee = 0.001 * E; /* keV */
  const double en = ee / m1;   /* keV/amu */
  double se;

  if (z1 == 1)
  {
    se = proton_stopping(e, z2, en); /* material.C:187-191 */
  }
  else if (z1 == 2)
  {
    /* material.C:192-212: He effective charge on top of the proton stopping */
    const double he0 = 1.0;
    double he = dmax(he0, en);
    double b = log(he);
    const double b2 = b * b;
    const double b4 = b2 * b2;
    double a = 0.2865 + 0.1266 * b - 0.001429 * b2 + 0.02402 * b * b2 - 0.01135 * b4 + 0.001475 * b4 * b;
    double heh = 1.0 - exp(-dmin(30.0, a));
    he = dmax(he, 1.0);
    const double t = 7.6 - log(he);
    a = 1.0 + (0.007 + 0.00005 * fz2) * exp(-(t * t));
    heh *= a * a;
    const double sp = proton_stopping(e, z2, he);
    se = sp * heh * 4.0;
    if (en <= he0)
      se *= sqrt(en / he0);
  }
  else
  {
    /* material.C:213-279: Brandt-Kitagawa heavy-ion scaling */
    const double yrmin = 0.13;
    double vrmin = 1.0;
    const double v = sqrt(en / 25.0) / vfermi;
    const double v2 = v * v;
    double vr;
    if (v >= 1.0)
      vr = v * vfermi * (1.0 + 1.0 / (5.0 * v2));
    else
      vr = (3.0 * vfermi / 4.0) * (1.0 + (2.0 * v2 / 3.0) - v2 * v2 / 15.0);

    const double cbrt_fz1 = cbrt(fz1);
    const double cbrt2_fz1 = cbrt_fz1 * cbrt_fz1;
    double yr = dmax(yrmin, vr / cbrt2_fz1);
    yr = dmax(yr, vrmin / cbrt2_fz1);
    const double yr03 = pow(yr, 0.3);
    double a = -0.803 * yr03 + 1.3167 * yr03 * yr03 + 0.38157 * yr + 0.008983 * yr * yr;

    /* ionisation level */
    const double q = dmin(1.0, dmax(0.0, 1.0 - exp(-dmin(a, 50.0))));

    const double b = (dmin(0.43, dmax(0.32, 0.12 + 0.025 * fz1))) / cbrt_fz1;
    const double l0 = (0.8 - q * dmin(1.2, 0.6 + fz1 / 30.0)) / cbrt_fz1;
    double l1;
    if (q < 0.2)
      l1 = 0.0;
    else if (q < dmax(0.0, 0.9 - 0.025 * fz1))
      l1 = b * (q - 0.2) / fabs(dmax(0.0, 0.9 - 0.025 * fz1) - 0.2000001);
    else if (q < dmax(0.0, 1.0 - 0.025 * dmin(16.0, fz1)))
      l1 = b;
    else
      l1 = b * (1.0 - q) / (0.025 * dmin(16.0, fz1));

    const double l = dmax(l1, l0 * lfctr);
    const double lx = 4.0 * l * vfermi / 1.919;
    double zeta = q + (1.0 / (2.0 * vfermi * vfermi)) * (1.0 - q) * log(1.0 + lx * lx);

    /* Z1^3 effect */
    const double t = 7.6 - dmax(0.0, log(en));
    a = -(t * t);
    zeta *= 1.0 + (1.0 / (fz1 * fz1)) * (0.18 + 0.0015 * fz2) * exp(a);

    if (yr <= dmax(yrmin, vrmin / cbrt2_fz1))
    {
      /* velocity-proportional stopping below yrmin */
      vrmin = dmax(vrmin, yrmin * cbrt2_fz1);
      const double vmin = 0.5 * (vrmin + sqrt(dmax(0.0, vrmin * vrmin - 0.8 * vfermi * vfermi)));
      const double eee = 25.0 * vmin * vmin;
      const double sp = proton_stopping(e, z2, eee);
      const double power = (z2 == 6 || ((z2 == 14 || z2 == 32) && z1 <= 19)) ? 0.375 : 0.5;
      const double zf = zeta * fz1;
      se = sp * (zf * zf) * pow(en / eee, power);
    }
    else
    {
      const double sp = proton_stopping(e, z2, en);
      const double zf = zeta * fz1;
      se = sp * (zf * zf);
    }
  }
  return se * 10.0;
}

/* MaterialBase::getrstop — material.C:113-122 */
static double
material_stopping(const orc_engine * e, const orc_material * M, int Z1, double m1, double E)
{
  double se = 0.0;
  for (int i = 0; i < M->n_elements; ++i)
    se += element_stopping(e, Z1, m1, E, M->el[i].Z) * M->el[i].t;
  return se * M->arho;
}

double
orc_getrstop(orc_engine * e, int material, int Z1, double m1, double E)
{
  if (material < 0 || material >= e->n_materials)
    return NAN;
  return material_stopping(e, &e->mat[material], Z1, m1, E);
}

int
orc_average(orc_engine * e, int material, int Z1, double m1, double * out, size_t capacity)
{
  if (material < 0 || material >= e->n_materials)
    return MTB_EINVAL;
  orc_material * M = &e->mat[material];
  if (capacity < (size_t)(6 + 4 * M->n_elements))
    return MTB_EINVAL;
  material_average(e, M, Z1, m1);
  out[0] = M->arho;
  out[1] = M->am;
  out[2] = M->az;
  out[3] = M->a;
  out[4] = M->f;
  out[5] = M->epsdg;
  for (int i = 0; i < M->n_elements; ++i)
  {
    out[6 + 4 * i + 0] = M->el[i].my;
    out[6 + 4 * i + 1] = M->el[i].ec;
    out[6 + 4 * i + 2] = M->el[i].ai;
    out[6 + 4 * i + 3] = M->el[i].fi;
  }
  M->averaged = 0;
  return MTB_OK;
}

/* ------------------------------------------------------------------------- */
/* geometry                                                                   */
/* ------------------------------------------------------------------------- */

/* sampleClusters::lookupCluster — sample_clusters.C:59-133.  Returns cluster index, -1 for the
 * matrix, -2 outside a CUT boundary. */
static int
clusters_lookup(const orc_engine * e, const double pos[3], double dr)
{
  const double * w = e->geom.w;
  int k[3], k1[3], k2[3], j[3];
  for (int i = 0; i < 3; ++i)
  {
    k[i] = (int)floor((pos[i] * e->kn[i]) / w[i]);
    if (pos[i] < 0.0 || pos[i] >= w[i])
    {
      switch (e->geom.bc[i])
      {
        case MTB_BC_CUT:
          return -2;
        case MTB_BC_INF:
          return -1;
        default:
          k[i] = k[i] % e->kn[i];
          if (k[i] < 0)
            k[i] += e->kn[i];
      }
    }
    const int ks = (int)((e->cmr + dr) / e->kd[i]) + 1;
    k1[i] = k[i] - ks;
    k2[i] = k[i] + ks;
    if (k1[i] < 0 && e->geom.bc[i] != MTB_BC_PBC)
      k1[i] = 0;
    if (k2[i] >= e->kn[i] && e->geom.bc[i] != MTB_BC_PBC)
      k2[i] = e->kn[i] - 1;
  }
  for (j[0] = k1[0]; j[0] <= k2[0]; ++j[0])
    for (j[1] = k1[1]; j[1] <= k2[1]; ++j[1])
      for (j[2] = k1[2]; j[2] <= k2[2]; ++j[2])
      {
        for (int i = 0; i < 3; ++i)
        {
          k[i] = j[i] % e->kn[i];
          if (k[i] < 0)
            k[i] += e->kn[i];
        }
        int l = e->sh[k[0] + e->kn[0] * (k[1] + e->kn[1] * k[2])];
        while (l >= 0)
        {
          double dif[3];
          for (int i = 0; i < 3; ++i)
          {
            dif[i] = pos[i] - e->c[i][l];
            if (e->geom.bc[i] == MTB_BC_PBC)
              dif[i] -= round(dif[i] / w[i]) * w[i];
          }
          double r2 = 0.0;
          for (int i = 0; i < 3; ++i)
            r2 += dif[i] * dif[i];
          const double rr = e->c[3][l] + dr;
          if (r2 < rr * rr)
            return l;
          l = e->cl[l];
        }
      }
  return -1;
}

/* The lookupMaterial() family.  Returns the material index or -1 for vacuum; *cluster receives
 * the cluster index the reference writes into material[1]->_tag (sample_clusters.C:53). */
static int
lookup_material(const orc_engine * e, const double pos[3], int * cluster)
{
  const mtb_geometry * g = &e->geom;
  *cluster = -1;
  switch (g->kind)
  {
    case MTB_GEOM_SOLID: /* sample_solid.C:25-29 */
      return 0;
    case MTB_GEOM_LAYERS: /* sample_layers.C:26-49 */
    {
      int i;
      double d = 0.0;
      for (i = 0; i < g->n_layers; ++i)
      {
        d += e->layer_thickness[i];
        if (pos[0] < d)
          break;
      }
      if (i >= e->n_materials)
        i = e->n_materials - 1;
      return i;
    }
    case MTB_GEOM_WIRE: /* sample_wire.C:37-46 */
    {
      const double x = (pos[0] / g->w[0]) * 2.0 - 1.0;
      const double y = (pos[1] / g->w[1]) * 2.0 - 1.0;
      return (x * x + y * y) > 1.0 ? -1 : 0;
    }
    case MTB_GEOM_BURIED_WIRE: /* sample_burried_wire.C:38-55 */
    {
      if (pos[2] < 0.0 && pos[2] >= -250.0)
        return 1;
      if (pos[2] > g->w[2] || pos[2] < -250.0)
        return -1;
      const double x = (pos[0] / g->w[0]) * 2.0 - 1.0;
      const double y = (pos[1] / g->w[1]) * 2.0 - 1.0;
      return (x * x + y * y) > 1.0 ? 1 : 0;
    }
    case MTB_GEOM_CLUSTERS: /* sample_clusters.C:43-55 */
    {
      const int l = clusters_lookup(e, pos, 0.0);
      if (l == -2)
        return -1;
      if (l == -1)
        return 0;
      *cluster = l;
      return 1;
    }
  }
  return -1;
}

int
orc_lookup_material(orc_engine * e, const double pos[3], int * cluster)
{
  int cl;
  const int m = lookup_material(e, pos, &cl);
  if (cluster)
    *cluster = cl;
  return m;
}

/* ------------------------------------------------------------------------- */
/* tallies (the five virtual hooks of trim.h:71-84 as tally modes)            */
/* ------------------------------------------------------------------------- */

static int
follow_recoil(orc_engine * e, const orc_ion * recoil, const orc_element * el)
{
  if (e->cfg.tally_mask & MTB_TALLY_PHONON)
    e->cnt.EnucTotal += el->Elbind; /* TrimPhononOut::followRecoil — trim.C:521-527 */
  switch (e->cfg.follow)
  {
    case MTB_FOLLOW_ALL:
      return 1;
    case MTB_FOLLOW_NONE:
      return 0;
    default:
      return recoil->gen < e->cfg.follow_max_gen; /* trim.h:120 */
  }
}

static void
vacancy_creation(orc_engine * e, const orc_ion * recoil, const orc_element * el, const orc_material * M)
{
  switch (e->cfg.vacancy_model)
  {
    case MTB_VAC_COUNT: /* trim.C:439-443 */
      e->vacancies_int++;
      break;
    case MTB_VAC_NRT: /* apps/src/TrimRange.C:31-47 */
    {
      const double Ed = el->Edisp;
      const double ed = 0.0115 * pow((double)recoil->Z, -7.0 / 3.0) * recoil->E;
      const double kd = 0.1337 * pow((double)recoil->Z, 2.0 / 3.0) / sqrt(recoil->m);
      const double g = 3.4008 * pow(ed, 1.0 / 6.0) + 0.40244 * pow(ed, 3.0 / 4.0) + ed;
      const double Ev = recoil->E / (1.0 + kd * g);
      if (Ev < Ed)
        break;
      if (Ev >= Ed / 0.4)
        e->vacancies_int = (int)((double)e->vacancies_int + Ev * 0.4 / Ed); /* int += double */
      else
        e->vacancies_int++;
      break;
    }
    case MTB_VAC_KP: /* TrimPrimaries::vacancyCreation — trim.C:445-464 */
      e->vacancies_int++;
      if (recoil->gen == e->cfg.follow_max_gen)
      {
        const double ed = 0.0115 * pow(M->az, -7.0 / 3.0) * recoil->E;
        const double g = 3.4008 * pow(ed, 1.0 / 6.0) + 0.40244 * pow(ed, 3.0 / 4.0) + ed;
        const double kd = 0.1337 * pow(M->az, 2.0 / 3.0) / sqrt(M->am);
        const double Ev = recoil->E / (1.0 + kd * g);
        e->vacancies_int += (int)(0.8 * Ev / (2.0 * el->Edisp));
      }
      break;
    default:
      break;
  }

  const int x = (int)recoil->pos[0]; /* truncation toward zero, TrimVacCount.C:35 */
  if ((e->cfg.tally_mask & MTB_TALLY_VAC_DEPTH) && x >= 0)
    hist_add(&e->vac, (size_t)x);
  if ((e->cfg.tally_mask & MTB_TALLY_VAC_ENERGY) && x >= 0)
  {
    int le = (int)log(recoil->E); /* TrimVacEnergyCount.C:43-44 */
    if (le < 0)
      le = 0;
    if ((size_t)le >= e->evac_rows)
    {
      e->evac = (orc_hist *)realloc(e->evac, sizeof(orc_hist) * (size_t)(le + 1));
      memset(e->evac + e->evac_rows, 0, sizeof(orc_hist) * ((size_t)le + 1 - e->evac_rows));
      e->evac_rows = (size_t)le + 1;
    }
    hist_add(&e->evac[le], (size_t)x);
  }
  if (e->cfg.tally_mask & MTB_TALLY_VACMAP) /* TrimVacMap::vacancyCreation — trim.C:483-501 */
  {
    int vx = (int)((recoil->pos[0] * MTB_VMAP_NX) / e->geom.w[0]);
    int vy = (int)((recoil->pos[1] * MTB_VMAP_NY) / e->geom.w[1]);
    vx -= (vx / MTB_VMAP_NX) * MTB_VMAP_NX;
    vy -= (vy / MTB_VMAP_NY) * MTB_VMAP_NY;
    int s = -1;
    if (recoil->Z == e->cfg.vmap_z[0])
      s = 0;
    else if (recoil->Z == e->cfg.vmap_z[1])
      s = 1;
    else if (recoil->Z == e->cfg.vmap_z[2])
      s = 2;
    if (s >= 0 && vx >= 0 && vy >= 0) /* the reference indexes out of bounds for negative bins */
      e->vmap[(vx * MTB_VMAP_NY + vy) * 3 + s]++;
  }
}

static void
replacement_collision(orc_engine * e, const orc_ion * recoil)
{
  e->cnt.replacements++;
  e->cas_repl++;
  if (e->cfg.tally_mask & MTB_TALLY_VAC_DEPTH) /* TrimVacCount.C:44-53 */
  {
    const int x = (int)recoil->pos[0];
    if (x >= 0)
      hist_add(&e->repl, (size_t)x);
  }
}

static void
dissipate_recoil_energy(orc_engine * e, const orc_ion * recoil, const orc_element * el)
{
  if (e->cfg.tally_mask & MTB_TALLY_RANGE) /* TrimRange::dissipateRecoilEnergy — TrimRange.C:49-54 */
  {
    if (e->range_n == e->range_cap)
    {
      e->range_cap = e->range_cap ? 2 * e->range_cap : 4096;
      e->range_x = (double *)realloc(e->range_x, sizeof(double) * e->range_cap);
      e->range_z = (int32_t *)realloc(e->range_z, sizeof(int32_t) * e->range_cap);
    }
    e->range_x[e->range_n] = recoil->pos[0];
    e->range_z[e->range_n] = recoil->Z;
    e->range_n++;
  }
  if (e->cfg.tally_mask & MTB_TALLY_PHONON) /* TrimPhononOut::dissipateRecoilEnergy — trim.C:513-519 */
    e->cnt.EnucTotal += recoil->E + el->Elbind;
}

static void
check_pka_state(orc_engine * e, const orc_ion * pka)
{
  if (e->cfg.tally_mask & MTB_TALLY_PHONON) /* TrimPhononOut::checkPKAState — trim.C:503-511 */
  {
    if (pka->state == MTB_MOVING || pka->state == MTB_LOST)
      return;
    e->cnt.EnucTotal += pka->E;
  }
}

/* ------------------------------------------------------------------------- */
/* the flight loop                                                            */
/* ------------------------------------------------------------------------- */

static inline double
vnorm3(const double v[3])
{
  return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
}

/* screened-potential value and derivative at reduced distance r — trim.C:194-227 */
static int
screened_potential(int pot, double r, double * v, double * v1)
{
  double ex1, ex2, ex3, ex4;
  switch (pot)
  {
    case MTB_POT_UNIVERSAL:
      ex1 = 0.18175 * exp(-3.1998 * r);
      ex2 = 0.50986 * exp(-0.94229 * r);
      ex3 = 0.28022 * exp(-0.4029 * r);
      ex4 = 0.028171 * exp(-0.20162 * r);
      *v = (ex1 + ex2 + ex3 + ex4) / r;
      *v1 = -(*v + 3.1998 * ex1 + 0.94229 * ex2 + 0.4029 * ex3 + 0.20162 * ex4) / r;
      return 0;
    case MTB_POT_MOLIERE:
      ex1 = exp(-0.3 * r);
      ex2 = (ex1 * ex1) * (ex1 * ex1);
      ex3 = ex2 * ((ex2 * ex2) * (ex2 * ex2)); /* Utility::pow<5>: x * pow<4>(x) */
      *v = (0.35 * ex1 + 0.55 * ex2 + 0.1 * ex3) / r;
      *v1 = -(*v + 0.105 * ex1 + 0.66 * ex2 + 0.6 * ex3) / r;
      return 0;
    case MTB_POT_CKR:
      ex1 = exp(-0.279 * r);
      ex2 = exp(-0.637 * r);
      ex3 = exp(-1.1919 * r);
      *v = (0.191 * ex1 + 0.474 * ex2 + 0.335 * ex3) / r;
      *v1 = -(*v + 0.531865 * ex1 + 0.30181 * ex2 + 0.6437 * ex3) / r;
      return 0;
  }
  return 1;
}

/* MAGIC fit parameters — trim.C:238-264 */
static void
magic_parameters(int pot, double eps, double sqe, double b, double * aa, double * ff)
{
  double cc;
  switch (pot)
  {
    case MTB_POT_MOLIERE:
      cc = (0.009611 + sqe) / (0.005175 + sqe);
      *aa = 2.0 * eps * (1.0 + (0.6743 / sqe)) * pow(b, cc);
      *ff = (sqrt(*aa * *aa + 1.0) - *aa) * ((6.314 + eps) / (10.0 + eps));
      break;
    case MTB_POT_CKR:
      cc = (0.235809 + sqe) / (0.126000 + sqe);
      *aa = 2.0 * eps * (1.0 + (1.0144 / sqe)) * pow(b, cc);
      *ff = (sqrt(*aa * *aa + 1.0) - *aa) * ((6935.0 + eps) / (83550.0 + eps));
      break;
    default:
      cc = (0.011615 + sqe) / (0.0071222 + sqe);
      *aa = 2.0 * eps * (1.0 + (0.99229 / sqe)) * pow(b, cc);
      *ff = (sqrt(*aa * *aa + 1.0) - *aa) * ((9.3066 + eps) / (14.813 + eps));
  }
}

/* TrimBase::trim — trim.C:35-425.  Follows one ion; survivors go to e->fifo. */
static void
transport_ion(orc_engine * e, orc_ion * pka, int collect_events)
{
  const double scale = e->cfg.length_scale;
  const double invscale = 1.0 / scale;
  const double tau = e->cfg.tau;
  const int pot = e->cfg.potential;
  const int mt = e->rng_mode == ORC_RNG_MT19937;

  pka->state = MTB_MOVING;
  e->cnt.ions++;
  for (int i = 0; i < e->n_materials; ++i)
    e->mat[i].averaged = 0; /* sample->averages(pka) before every trim() — runmytrim.C:85 */

  double r1 = 0.0;
  if (mt)
    r1 = orc_mt_drand(&e->mt); /* trim.C:71 */

  do
  {
    ++pka->ic;

    int cluster;
    const int mi = lookup_material(e, pka->pos, &cluster); /* trim.C:80 */
    if (mi < 0)
    {
      e->cnt.left_sample++;
      break; /* vacuum: state stays MOVING */
    }
    orc_material * M = &e->mat[mi];
    if (!M->averaged)
      material_average(e, M, pka->Z, pka->m);
    const int mtag = (e->geom.kind == MTB_GEOM_CLUSTERS && mi == 1) ? cluster : M->tag;
    if (e->geom.kind == MTB_GEOM_CLUSTERS && mi == 1)
      M->tag = cluster; /* sticky side effect of sample_clusters.C:53 */

    e->cnt.steps++;

    /* v_norm(dir) — trim.C:85, functions.h:66-70 */
    {
      const double s = 1.0 / vnorm3(pka->dir);
      pka->dir[0] *= s;
      pka->dir[1] *= s;
      pka->dir[2] *= s;
    }

    /* random numbers of this step */
    double r2, hh, uphi = 0.0;
    uint32_t w3 = 0;
    if (mt)
    {
      r2 = orc_mt_drand(&e->mt); /* trim.C:143 */
      hh = orc_mt_drand(&e->mt); /* trim.C:147 */
    }
    else
    {
      const uint32_t ctr[4] = {pka->ic, (uint32_t)pka->uid, (uint32_t)(pka->uid >> 32), 0u};
      uint32_t w[4];
      orc_philox4x32_10(ctr, e->key, w);
      r2 = (double)orc_u01(w[0]);
      hh = (double)orc_u01(w[1]);
      uphi = (double)orc_u01(w[2]);
      r1 = (double)orc_u01(w[3]);
      w3 = w[3];
    }

    /* maximum impact parameter and free flight path — trim.C:88-94 */
    double eps = pka->E * M->f;
    const double eeg = sqrt(eps * M->epsdg);
    M->pmax = M->a / (eeg + sqrt(eeg) + 0.125 * pow(eeg, 0.1));
    double ls = 1.0 / (M_PI * (M->pmax * M->pmax) * M->arho);
    if (pka->ic == 1)
      ls = r1 * dmin(ls, e->cfg.cw);

    /* impact parameter — trim.C:143-144 */
    const double p = M->pmax * sqrt(r2);

    /* target element — trim.C:147-156 */
    int nn;
    for (nn = 0; nn < M->n_elements; ++nn)
    {
      hh -= M->el[nn].t;
      if (hh <= 0)
        break;
    }
    if (nn >= M->n_elements)
      nn = M->n_elements - 1; /* the reference reads past the end here (measure-zero event) */
    const orc_element * el = &M->el[nn];

    eps = el->fi * pka->E; /* trim.C:159-160 */
    const double b = p / el->ai;

    const double see = material_stopping(e, M, pka->Z, pka->m, pka->E); /* trim.C:166 */
    double dee = ls * see;

    double s2, c2, ct, st;
    if (eps > 10.0)
    {
      /* Rutherford — trim.C:172-179 */
      const double t = 2.0 * eps * b;
      s2 = 1.0 / (1.0 + (1.0 + b * (1.0 + b)) * (t * t));
      c2 = 1.0 - s2;
      ct = 2.0 * c2 - 1.0;
      st = sqrt(1.0 - ct * ct);
    }
    else
    {
      /* closest approach by Newton iteration — trim.C:182-233 */
      double r = b;
      double rr = -2.7 * log(eps * b);
      if (rr >= b)
      {
        rr = -2.7 * log(eps * rr);
        if (rr >= b)
          r = rr;
      }
      double v = 0.0, v1 = 0.0, q;
      do
      {
        screened_potential(pot, r, &v, &v1);
        const double fr = b * b / r + v * r / eps - r;
        const double fr1 = -b * b / (r * r) + (v + v1 * r) / eps - 1.0;
        q = fr / fr1;
        r -= q;
      } while (fabs(q / r) > 0.001);

      const double roc = -2.0 * (eps - v) / v1;
      const double sqe = sqrt(eps);
      double aa, ff;
      magic_parameters(pot, eps, sqe, b, &aa, &ff);
      const double delta = (r - b) * aa * ff / (ff + 1.0);
      const double co = (b + delta + roc) / (r + roc);
      c2 = co * co;
      s2 = 1.0 - c2;
      ct = 2.0 * c2 - 1.0;
      st = sqrt(1.0 - ct * ct);
    }

    /* energy transfer, electronic loss — trim.C:275-296 */
    double den = el->ec * s2 * pka->E;
    if (dee > pka->E)
      dee = pka->E;
    pka->E -= dee;
    e->cnt.EelTotal += dee;

    const double p1 = sqrt(2.0 * pka->m * pka->E);
    if (den > pka->E)
      den = pka->E;
    pka->E -= den;
    const double p2 = sqrt(2.0 * pka->m * pka->E);

    /* recoil is born at the previous collision site — trim.C:306-318, ion.C:54-68 */
    orc_ion rec;
    memset(&rec, 0, sizeof(rec));
    rec.gen = pka->gen + 1;
    rec.Ef = pka->Ef;
    rec.tag = -1;
    rec.state = MTB_MOVING;
    rec.primary = pka->primary;
    for (int i = 0; i < 3; ++i)
    {
      rec.pos[i] = pka->pos[i];
      pka->pos[i] += pka->dir[i] * (ls - tau) * invscale;
      rec.dir[i] = pka->dir[i] * p1;
    }
    rec.E = den;
    rec.E -= el->Elbind;
    rec.m = el->m;
    rec.Z = el->Z;

    /* random unit vector perpendicular to dir */
    double perp[3];
    if (mt)
    {
      /* trim.C:322-333 */
      double norm;
      do
      {
        double rd[3];
        do
        {
          for (int i = 0; i < 3; ++i)
            rd[i] = 2.0 * orc_mt_drand(&e->mt) - 1.0;
        } while (rd[0] * rd[0] + rd[1] * rd[1] + rd[2] * rd[2] > 1.0);
        for (int i = 0; i < 3; ++i)
          perp[i] = pka->dir[(i + 1) % 3] * rd[(i + 2) % 3] - pka->dir[(i + 2) % 3] * rd[(i + 1) % 3];
        norm = vnorm3(perp);
      } while (norm == 0.0);
      for (int i = 0; i < 3; ++i)
        perp[i] /= norm;
    }
    else
    {
      /* dir x (uniform ball vector), normalised, is uniform on the circle perpendicular to dir;
       * draw its azimuth directly in a branch-free orthonormal frame (Duff et al. 2017). */
      const double * d = pka->dir;
      const double sg = copysign(1.0, d[2]);
      const double a = -1.0 / (sg + d[2]);
      const double bb = d[0] * d[1] * a;
      const double e1[3] = {1.0 + sg * d[0] * d[0] * a, sg * bb, -sg * d[0]};
      const double e2[3] = {bb, sg + d[1] * d[1] * a, -d[1]};
      const double phi = 2.0 * M_PI * uphi;
      const double cp = cos(phi), sp = sin(phi);
      for (int i = 0; i < 3; ++i)
        perp[i] = cp * e1[i] + sp * e2[i];
    }

    /* lab scattering angle and new directions — trim.C:336-341 */
    const double psi = atan2(st, ct + el->my);
    const double cpsi = cos(psi), spsi = sin(psi);
    for (int i = 0; i < 3; ++i)
    {
      pka->dir[i] *= cpsi;
      pka->dir[i] += perp[i] * spsi;
      rec.dir[i] -= pka->dir[i] * p2;
    }

    /* CUT boundaries — trim.C:344-352 */
    for (int i = 0; i < 3; ++i)
      if (e->geom.bc[i] == MTB_BC_CUT && (pka->pos[i] > e->geom.w[i] || pka->pos[i] < 0.0))
      {
        pka->state = MTB_LOST;
        e->cnt.lost++;
        break;
      }

    /* fate of recoil and pka — trim.C:357-411 */
    int above = 0, followed = 0;
    if (pka->state != MTB_LOST)
    {
      if (rec.E > el->Edisp - el->Elbind)
      {
        above = 1;
        if (follow_recoil(e, &rec, el))
        {
          const double s = 1.0 / vnorm3(rec.dir);
          rec.dir[0] *= s;
          rec.dir[1] *= s;
          rec.dir[2] *= s;
          rec.tag = mtag;
          rec.id = e->next_id++;
          rec.uid = orc_child_uid(pka->uid, pka->ic, w3);
          rec.ic = 0;
          memcpy(rec.pos0, rec.pos, sizeof(rec.pos0));
          rec.E0 = rec.E;
          followed = 1;
        }
        else
          rec.state = MTB_DELETE;

        if (pka->E > el->Edisp)
          vacancy_creation(e, &rec, el, M);
        else
        {
          replacement_collision(e, &rec);
          pka->state = pka->Z == el->Z ? MTB_REPLACEMENT : MTB_SUBSTITUTIONAL;
        }
      }
      else
      {
        dissipate_recoil_energy(e, &rec, el);
        rec.state = MTB_DELETE;
        if (pka->E < pka->Ef)
          pka->state = MTB_INTERSTITIAL;
      }
    }

    if (collect_events)
    {
      if (e->ev_n < e->ev_cap)
      {
        mtb_event * ev = &e->ev[e->ev_n];
        memset(ev, 0, sizeof(*ev));
        for (int i = 0; i < 3; ++i)
        {
          ev->pka_pos[i] = pka->pos[i];
          ev->pka_dir[i] = pka->dir[i];
          ev->recoil_pos[i] = rec.pos[i];
          ev->recoil_dir[i] = rec.dir[i];
        }
        if (above && !followed)
        {
          const double s = 1.0 / vnorm3(rec.dir);
          for (int i = 0; i < 3; ++i)
            ev->recoil_dir[i] = rec.dir[i] * s;
        }
        ev->pka_E = pka->E;
        ev->recoil_E = rec.E;
        ev->ls = ls;
        ev->dee = dee;
        ev->den = den;
        ev->material = mi;
        ev->element = nn;
        ev->material_tag = mtag;
        ev->pka_state = pka->state;
        ev->recoil_above_threshold = above;
      }
      e->ev_n++;
    }
    else if (followed)
    {
      e->cnt.recoils_queued++;
      fifo_push(&e->fifo, &rec);
    }

    check_pka_state(e, pka); /* trim.C:418 */

  } while (pka->state == MTB_MOVING);
}

static void
log_ion(orc_engine * e, const orc_ion * ion)
{
  if (!(e->cfg.tally_mask & MTB_TALLY_IONLOG))
    return;
  if (e->cfg.ionlog_z && ion->Z != e->cfg.ionlog_z)
    return;
  if (e->ilog_n == e->ilog_cap)
  {
    e->ilog_cap = e->ilog_cap ? 2 * e->ilog_cap : 1024;
    e->ilog = (mtb_ion_log *)realloc(e->ilog, sizeof(mtb_ion_log) * e->ilog_cap);
  }
  mtb_ion_log * L = &e->ilog[e->ilog_n++];
  memset(L, 0, sizeof(*L));
  memcpy(L->pos0, ion->pos0, sizeof(L->pos0));
  memcpy(L->pos1, ion->pos, sizeof(L->pos1));
  L->E0 = ion->E0;
  L->E1 = ion->E;
  L->uid = ion->uid;
  L->primary = ion->primary;
  L->Z = ion->Z;
  L->gen = ion->gen;
  L->tag = ion->tag;
  L->state = ion->state;
}

static void
ion_from_abi(orc_ion * o, const mtb_ion * in)
{
  memset(o, 0, sizeof(*o));
  memcpy(o->pos, in->pos, sizeof(o->pos));
  memcpy(o->dir, in->dir, sizeof(o->dir));
  o->E = in->E;
  o->m = in->m;
  o->Ef = in->Ef;
  o->Z = in->Z;
  o->gen = in->gen;
  o->tag = in->tag;
  o->state = MTB_MOVING;
  memcpy(o->pos0, in->pos, sizeof(o->pos0));
  o->E0 = in->E;
}

/* The per-primary FIFO loop — apps/runmytrim.C:76-92 */
int
orc_run(orc_engine * e, uint64_t n, const mtb_ion * primaries, uint64_t seed, uint64_t first_index,
        mtb_record * records)
{
  if (!e->n_materials)
    return MTB_EINVAL;
  e->key[0] = (uint32_t)seed;
  e->key[1] = (uint32_t)(seed >> 32);
  for (uint64_t ip = 0; ip < n; ++ip)
  {
    orc_ion pka;
    ion_from_abi(&pka, &primaries[ip]);
    pka.uid = first_index + ip;
    pka.primary = first_index + ip;
    pka.is_primary = 1;
    pka.id = e->next_id++;
    if (e->rng_mode == ORC_RNG_MT19937)
      orc_mt_seed(&e->mt, primaries[ip].seed); /* runmytrim.C:78 */

    const mtb_counters before = e->cnt;
    const int vac_before = e->vacancies_int;
    e->cas_repl = 0;
    mtb_record rec;
    memset(&rec, 0, sizeof(rec));

    fifo_push(&e->fifo, &pka);
    orc_ion ion;
    while (fifo_pop(&e->fifo, &ion))
    {
      const uint64_t steps0 = e->cnt.steps;
      transport_ion(e, &ion, 0);
      log_ion(e, &ion);
      if (ion.is_primary)
      {
        memcpy(rec.pos, ion.pos, sizeof(rec.pos));
        rec.E = ion.E;
        rec.state = ion.state;
        rec.primary_steps = (uint32_t)(e->cnt.steps - steps0);
      }
    }
    e->cnt.primaries++;
    if (records)
    {
      rec.Eel = e->cnt.EelTotal - before.EelTotal;
      rec.Enuc = e->cnt.EnucTotal - before.EnucTotal;
      rec.vacancies = (uint32_t)(e->vacancies_int - vac_before);
      rec.replacements = e->cas_repl;
      rec.steps = (uint32_t)(e->cnt.steps - before.steps);
      rec.ions = (uint32_t)(e->cnt.ions - before.ions);
      records[ip] = rec;
    }
  }
  e->cnt.vacancies_created = (uint64_t)(int64_t)e->vacancies_int;
  return MTB_OK;
}

int
orc_trim_one(orc_engine * e, mtb_ion * ion, uint64_t seed, uint64_t uid, int32_t * final_state,
             mtb_event * events, size_t capacity, size_t * n_events)
{
  if (!e->n_materials)
    return MTB_EINVAL;
  e->key[0] = (uint32_t)seed;
  e->key[1] = (uint32_t)(seed >> 32);
  orc_ion pka;
  ion_from_abi(&pka, ion);
  pka.uid = uid;
  pka.primary = uid;
  if (e->rng_mode == ORC_RNG_MT19937)
    orc_mt_seed(&e->mt, ion->seed);
  e->ev = events;
  e->ev_cap = events ? capacity : 0;
  e->ev_n = 0;
  transport_ion(e, &pka, 1);
  e->cnt.vacancies_created = (uint64_t)(int64_t)e->vacancies_int;
  memcpy(ion->pos, pka.pos, sizeof(pka.pos));
  memcpy(ion->dir, pka.dir, sizeof(pka.dir));
  ion->E = pka.E;
  if (final_state)
    *final_state = pka.state;
  if (n_events)
    *n_events = e->ev_n;
  e->ev = 0;
  return e->ev_n > e->ev_cap && events ? MTB_ECAPACITY : MTB_OK;
}

/* ------------------------------------------------------------------------- */
/* tally read-back                                                            */
/* ------------------------------------------------------------------------- */

int
orc_get_counters(orc_engine * e, mtb_counters * out)
{
  *out = e->cnt;
  out->vacancies_created = (uint64_t)(int64_t)e->vacancies_int;
  return MTB_OK;
}

int
orc_get_vac_depth(orc_engine * e, uint64_t * vac, uint64_t * repl, size_t capacity, size_t * n_bins)
{
  const size_t n = e->vac.n > e->repl.n ? e->vac.n : e->repl.n; /* TrimVacCount.C:72-74 */
  if (n_bins)
    *n_bins = n;
  for (size_t i = 0; i < capacity; ++i)
  {
    if (vac)
      vac[i] = i < e->vac.n ? e->vac.v[i] : 0;
    if (repl)
      repl[i] = i < e->repl.n ? e->repl.v[i] : 0;
  }
  return n > capacity ? MTB_ECAPACITY : MTB_OK;
}

int
orc_get_vac_energy(orc_engine * e, uint64_t * evac, size_t rows, size_t bins)
{
  memset(evac, 0, rows * bins * sizeof(uint64_t));
  for (size_t r = 0; r < e->evac_rows && r < rows; ++r)
    for (size_t x = 0; x < e->evac[r].n && x < bins; ++x)
      evac[r * bins + x] = e->evac[r].v[x];
  return MTB_OK;
}

int
orc_get_vacmap(orc_engine * e, uint64_t * vmap)
{
  memcpy(vmap, e->vmap, sizeof(e->vmap));
  return MTB_OK;
}

int
orc_get_range_list(orc_engine * e, double * x, int32_t * Z, size_t capacity, size_t * n)
{
  if (n)
    *n = e->range_n;
  for (size_t i = 0; i < e->range_n && i < capacity; ++i)
  {
    if (x)
      x[i] = e->range_x[i];
    if (Z)
      Z[i] = e->range_z[i];
  }
  return e->range_n > capacity ? MTB_ECAPACITY : MTB_OK;
}

int
orc_get_ion_log(orc_engine * e, mtb_ion_log * out, size_t capacity, size_t * n)
{
  if (n)
    *n = e->ilog_n;
  for (size_t i = 0; i < e->ilog_n && i < capacity; ++i)
    out[i] = e->ilog[i];
  return e->ilog_n > capacity ? MTB_ECAPACITY : MTB_OK;
}

/* ------------------------------------------------------------------------- */
/* fission source + the mytrim_uo2 experiment (gold-file pin)                 */
/* ------------------------------------------------------------------------- */

/* MassInverter::f — invert.C:47-56 (single-precision erff, as in the reference) */
static double
fission_mass_cdf(double x)
{
  return (100.088 + 0.112798 * erff(-5.56257 + 0.0471405 * x) + 37.4781 * erff(-19.3772 + 0.137386 * x) +
          37.4781 * erff(-13.0462 + 0.137386 * x) + 12.5094 * erff(-30.8853 + 0.229537 * x) +
          12.5094 * erff(-23.2853 + 0.229537 * x)) /
         200.1756;
}

/* EnergyInverter::f — invert.C:58-63 */
static double
fission_energy_cdf(double A, double x)
{
  const double x1 = x / (1.0 - A / 234.0);
  return (-0.00014122 + (0.00014122 - 7.12299E-7 * x1) * exp(0.0886603 * x1)) / 127.216;
}

/* Inverter::x — invert.C:25-45: 32-step bisection of f(x)/f(maxx) */
double
orc_mass_inverter_x(double f1)
{
  const double maxx = 235.0, tol = 1e-7, maxf = fission_mass_cdf(maxx);
  double x1 = maxx / 2.0, w = maxx / 4.0;
  for (int i = 0; i < 32; ++i)
  {
    const double f2 = fission_mass_cdf(x1) / maxf;
    if (fabs(f2 - f1) <= tol)
      break;
    if (f2 > f1)
      x1 -= w;
    else
      x1 += w;
    w *= 0.5;
  }
  return x1;
}

double
orc_energy_inverter_x(double A, double f1)
{
  const double maxx = 186.98, tol = 1e-7, maxf = fission_energy_cdf(A, maxx);
  double x1 = maxx / 2.0, w = maxx / 4.0;
  for (int i = 0; i < 32; ++i)
  {
    const double f2 = fission_energy_cdf(A, x1) / maxf;
    if (fabs(f2 - f1) <= tol)
      break;
    if (f2 > f1)
      x1 -= w;
    else
      x1 += w;
    w *= 0.5;
  }
  return x1;
}

/* apps/mytrim_uo2.C:49-358 with TrimBase (mode PLAIN), one shared mt19937 stream. */
int
orc_uo2_experiment(const char * base, double r, double Cbf, int Nev, uint32_t seed, double * Eel_out,
                   double * Efiss_out)
{
  mtb_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.tmin = 0.2;
  cfg.tau = 0.0;
  cfg.cw = 0.001;
  cfg.length_scale = 1.0;
  cfg.follow = MTB_FOLLOW_ALL;
  cfg.vacancy_model = MTB_VAC_COUNT;
  orc_engine * e = orc_create(&cfg, ORC_RNG_MT19937);
  orc_mt_seed(&e->mt, seed);

  /* sample + clusters (mytrim_uo2.C:95-142) */
  e->geom.kind = MTB_GEOM_CLUSTERS;
  for (int i = 0; i < 3; ++i)
  {
    e->geom.w[i] = 400.0;
    e->geom.bc[i] = MTB_BC_PBC;
  }
  const int kn = (int)(400.0 / r) - 1;
  clusters_init_hash(e, kn, kn, kn);
  const double v_sam = e->geom.w[0] * e->geom.w[1] * e->geom.w[2];
  const int n_cl = (int)(v_sam * 7.0e-7 * Cbf);
  /* addRandomClusters(n_cl, r, 25.0) — sample_clusters.C:227-252 */
  for (int i = 0; i < n_cl; ++i)
    for (;;)
    {
      double npos[3];
      for (int j = 0; j < 3; ++j)
        npos[j] = orc_mt_drand(&e->mt) * e->geom.w[j];
      if (clusters_lookup(e, npos, 25.0 + r) == -1)
      {
        clusters_add(e, npos[0], npos[1], npos[2], r);
        break;
      }
    }

  char fname[512];
  snprintf(fname, sizeof(fname), "%s.clcoor", base);
  FILE * ccf = fopen(fname, "wt");
  if (!ccf)
    return MTB_EINVAL;
  for (int i = 0; i < e->cn; ++i)
    fprintf(ccf, "%f %f %f %f %d\n", e->c[0][i], e->c[1][i], e->c[2][i], e->c[3][i], i);
  fclose(ccf);

  /* materials (mytrim_uo2.C:163-184) */
  const mtb_element els[3] = {{92, 0, 235.0, 1.0, 25.0, 3.0}, {8, 0, 16.0, 2.0, 25.0, 3.0}, {54, 0, 132.0, 1.0, 25.0, 3.0}};
  const mtb_material mats[2] = {{10.0, -1, 2, 0, 0}, {3.5, -1, 1, 2, 0}};
  orc_set_materials(e, 2, mats, 3, els);
  const int gas_z1 = 54;

  snprintf(fname, sizeof(fname), "%s.Erec", base);
  FILE * erec = fopen(fname, "wt");
  snprintf(fname, sizeof(fname), "%s.dist", base);
  FILE * rdist = fopen(fname, "wt");
  if (!erec || !rdist)
    return MTB_EINVAL;

  double Eel_sum = 0.0, Efiss_sum = 0.0;
  for (int n = 0; n < Nev; ++n)
  {
    /* fission fragment pair (mytrim_uo2.C:226-266) */
    const double A1 = orc_mass_inverter_x(orc_mt_drand(&e->mt));
    const double A2 = 235.0 - A1;
    const double Etot = orc_energy_inverter_x(A1, orc_mt_drand(&e->mt));
    const double E1 = Etot * A2 / (A1 + A2);
    const double E2 = Etot - E1;
    const int Z1 = (int)round((A1 * 92.0) / 235.0);
    const int Z2 = 92 - Z1;

    orc_ion ff1;
    memset(&ff1, 0, sizeof(ff1));
    ff1.gen = 0;
    ff1.tag = -1;
    ff1.Z = Z1;
    ff1.m = A1;
    ff1.E = E1 * 1.0e6;
    ff1.Ef = 3.0;
    double norm;
    do
    {
      for (int i = 0; i < 3; ++i)
        ff1.dir[i] = 2.0 * orc_mt_drand(&e->mt) - 1.0;
      norm = ff1.dir[0] * ff1.dir[0] + ff1.dir[1] * ff1.dir[1] + ff1.dir[2] * ff1.dir[2];
    } while (norm <= 0.0001 || norm > 1.0);
    {
      const double s = sqrt(norm);
      for (int i = 0; i < 3; ++i)
        ff1.dir[i] /= s;
    }
    for (int i = 0; i < 3; ++i)
      ff1.pos[i] = orc_mt_drand(&e->mt) * e->geom.w[i];
    ff1.state = MTB_MOVING;
    orc_ion ff2 = ff1;
    for (int i = 0; i < 3; ++i)
      ff2.dir[i] = -ff2.dir[i];
    ff2.Z = Z2;
    ff2.m = A2;
    ff2.E = E2 * 1.0e6;
    fifo_push(&e->fifo, &ff1);
    fifo_push(&e->fifo, &ff2);
    Efiss_sum += ff1.E + ff2.E;

    orc_ion pka;
    double pos1[3] = {0, 0, 0};
    while (fifo_pop(&e->fifo, &pka))
    {
      int md = 0;
      if (pka.Z == gas_z1)
      {
        /* pre-cascade analysis (mytrim_uo2.C:281-308) */
        if (pka.E > 200 && pka.E < 12000)
          md = 1;
        if (pka.gen > 0)
          fprintf(erec, "%f\t%d\t%d\n", pka.E, pka.gen, md);
        if (pka.tag >= 0)
          for (int i = 0; i < 3; ++i)
          {
            double dif = e->c[i][pka.tag] - pka.pos[i];
            if (e->geom.bc[i] == MTB_BC_PBC)
              dif -= round(dif / e->geom.w[i]) * e->geom.w[i];
            pos1[i] = pka.pos[i] + dif;
          }
      }
      transport_ion(e, &pka, 0);
      if (pka.Z == gas_z1 && pka.tag >= 0)
      {
        /* post-cascade analysis (mytrim_uo2.C:319-338) */
        double dif[3], d2 = 0.0;
        for (int i = 0; i < 3; ++i)
        {
          dif[i] = pos1[i] - pka.pos[i];
          d2 += dif[i] * dif[i];
        }
        fprintf(rdist, "%f %d %f %f %f\n", sqrt(d2), md, pka.pos[0], pka.pos[1], pka.pos[2]);
      }
    }
    Eel_sum += e->cnt.EelTotal;
    e->cnt.EelTotal = 0.0;
  }
  fclose(rdist);
  fclose(erec);
  if (Eel_out)
    *Eel_out = Eel_sum;
  if (Efiss_out)
    *Efiss_out = Efiss_sum;
  orc_destroy(e);
  return MTB_OK;
}
