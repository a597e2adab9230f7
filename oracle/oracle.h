/*
 * oracle.h — CPU restatement of MyTRIM's cascade-transport hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under mytrim_b200/ or include/ may call,
 * link or include this; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg use it, and only as the checker.
 *
 * What it is: a plain-C, double-precision restatement of the reference
 * algorithm (TrimBase::trim, MaterialBase::{prepare,average,getrstop,rstop,
 * rpstop}, the Sample*::lookupMaterial family, the in-tree tally hooks and the
 * per-primary FIFO loop), every function citing the reference file:line it
 * follows.  It takes the same plain-struct configuration as the C ABI
 * (include/mytrim_b200.h) so tests feed both sides identical inputs.
 *
 * Two random-number modes:
 *   ORC_RNG_MT19937 — one std::mt19937 stream per primary, consumed in FIFO
 *     order exactly as the reference does (simconf.h:52, runmytrim.C:76-92),
 *     including the 3-D rejection loop of trim.C:322-332.  In this mode the
 *     oracle reproduces the compiled reference bit-for-bit; that is how it is
 *     pinned (tests/test_oracle_vs_reference.py, tests/golden/).
 *   ORC_RNG_PHILOX — the scheduling-independent per-ion Philox4x32-10 streams
 *     the GPU kernel uses (one block per collision step, azimuth drawn
 *     directly).  Same physics, same arithmetic order; only the source of
 *     uniforms differs.  This is the deterministic partner of the CUDA path.
 *
 * Parity status: PINNED — see DESIGN.md §oracle.
 */
#ifndef MYTRIM_ORACLE_H
#define MYTRIM_ORACLE_H

#include "../include/mytrim_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_RNG_MT19937 = 0, ORC_RNG_PHILOX = 1 };

typedef struct orc_engine orc_engine;

orc_engine * orc_create(const mtb_config * cfg, int rng_mode);
void orc_destroy(orc_engine * e);

int orc_set_tables(orc_engine * e, const double * pcoef, const double * vfermi, const double * lfctr,
                   const double * mm1);
int orc_set_materials(orc_engine * e, int n_materials, const mtb_material * materials, int n_elements,
                      const mtb_element * elements);
int orc_set_geometry(orc_engine * e, const mtb_geometry * geom);

/* n primaries; in MT mode primary i reseeds the stream with primaries[i].seed. */
int orc_run(orc_engine * e, uint64_t n, const mtb_ion * primaries, uint64_t seed, uint64_t first_index,
            mtb_record * records);

int orc_reset_tallies(orc_engine * e);
int orc_get_counters(orc_engine * e, mtb_counters * out);
int orc_get_vac_depth(orc_engine * e, uint64_t * vac, uint64_t * repl, size_t capacity, size_t * n_bins);
int orc_get_vac_energy(orc_engine * e, uint64_t * evac, size_t rows, size_t bins);
int orc_get_vacmap(orc_engine * e, uint64_t * vmap);
int orc_get_range_list(orc_engine * e, double * x, int32_t * Z, size_t capacity, size_t * n);
int orc_get_ion_log(orc_engine * e, mtb_ion_log * out, size_t capacity, size_t * n);

/* follows ONE ion without following recoils, reporting every collision (mirror of mtb_trim_one) */
int orc_trim_one(orc_engine * e, mtb_ion * ion, uint64_t seed, uint64_t uid, int32_t * final_state,
                 mtb_event * events, size_t capacity, size_t * n_events);

/* RNG-free pieces, exposed for known-answer tests */
double orc_getrstop(orc_engine * e, int material, int Z1, double m1, double E);
/* out: arho, am, az, a, f, epsdg, then per element my, ec, ai, fi */
int orc_average(orc_engine * e, int material, int Z1, double m1, double * out, size_t capacity);
int orc_lookup_material(orc_engine * e, const double pos[3], int * cluster);

/* random number primitives */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
uint64_t orc_child_uid(uint64_t uid, uint32_t ic, uint32_t w3);
float orc_u01(uint32_t x);
typedef struct { uint32_t mt[624]; int idx; } orc_mt19937;
void orc_mt_seed(orc_mt19937 * g, uint32_t seed);
uint32_t orc_mt_next(orc_mt19937 * g);
double orc_mt_drand(orc_mt19937 * g);   /* std::uniform_real_distribution<double>(0,1) */
uint32_t orc_mt_irand(orc_mt19937 * g); /* std::uniform_int_distribution<unsigned>(0,65535) */

/* fission source + cluster placement of apps/mytrim_uo2.C, for the gold-file test */
double orc_mass_inverter_x(double f);
double orc_energy_inverter_x(double A, double f);
/* Runs the `mytrim_uo2 base r Cbf Nev` experiment with MYTRIM_SEED=seed in MT mode and
 * writes base.Erec / base.clcoor / base.dist in the reference's formats. */
int orc_uo2_experiment(const char * base, double r, double Cbf, int Nev, uint32_t seed,
                       double * Eel_out, double * Efiss_out);

#ifdef __cplusplus
}
#endif
#endif
