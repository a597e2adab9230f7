/*
 * mytrim_b200.h — C ABI of the B200-native cascade-transport engine.
 *
 * This is the drop-in boundary for MyTRIM's hot path: the TrimBase::trim()
 * flight loop (reference trim.C:35-425) together with the per-primary recoil
 * queue loop every MyTRIM app wraps around it (apps/runmytrim.C:71-93,
 * apps/mytrim_uo2.C:274-342, apps/mytrim_layers.C:156-194).  The reference has
 * no FFI: its plugin surface is C++ inheritance.  The C++
 * façade in include/mytrim/ re-creates that surface and flattens it into the
 * plain structs below; any other host language can bind these entry points
 * directly.
 *
 * Conventions: every entry point returns an mtb_status (never exits, never
 * throws); all pointers are HOST pointers unless the name says `_dev`; buffers
 * are caller-owned; one handle per host thread / per GPU (a handle is not
 * re-entrant, like a reference TrimBase instance).  Units as in the reference:
 * eV, Angstrom (divided by length_scale for positions), amu, g/cm^3.
 */
#ifndef MYTRIM_B200_H
#define MYTRIM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTB_NZ 92            /* rows of the ZBL tables, Z = 1..92 (simconf.h:98) */
#define MTB_MAX_RANGE_Z 112  /* TrimRange reserves 112 Z slots (apps/src/TrimRange.C:27) */
#define MTB_VMAP_NX 20       /* TrimVacMap grid (trim.h:182) */
#define MTB_VMAP_NY 20

typedef enum {
  MTB_OK = 0,
  MTB_EINVAL = 1,    /* bad argument / inconsistent configuration */
  MTB_ECUDA = 2,     /* CUDA runtime error, see mtb_last_error() */
  MTB_ENODEV = 3,    /* no usable CUDA device */
  MTB_ESTACK = 4,    /* a cascade overflowed the per-lane recoil stack */
  MTB_ENOMEM = 5,
  MTB_ECAPACITY = 6, /* an output list (ion log, range list, events) overflowed; counts are still exact */
  MTB_ENCCL = 7
} mtb_status;

/* trim.h:63-69 */
typedef enum { MTB_POT_UNIVERSAL = 0, MTB_POT_MOLIERE = 1, MTB_POT_CKR = 2 } mtb_potential;
/* sample.h:49-54 */
typedef enum { MTB_BC_PBC = 0, MTB_BC_INF = 1, MTB_BC_CUT = 2 } mtb_boundary;
/* ion.h:76-85 */
typedef enum {
  MTB_MOVING = 0, MTB_REPLACEMENT = 1, MTB_SUBSTITUTIONAL = 2, MTB_INTERSTITIAL = 3,
  MTB_LOST = 4, MTB_DELETE = 5, MTB_VACANCY = 6
} mtb_ion_state;

/* Which SampleBase subclass the geometry lookup reproduces. */
typedef enum {
  MTB_GEOM_SOLID = 0,       /* sample_solid.C:25-29 */
  MTB_GEOM_LAYERS = 1,      /* sample_layers.C:26-49 */
  MTB_GEOM_WIRE = 2,        /* sample_wire.C:37-46 */
  MTB_GEOM_BURIED_WIRE = 3, /* sample_burried_wire.C:38-55 */
  MTB_GEOM_CLUSTERS = 4     /* sample_clusters.C:43-133 */
} mtb_geometry_kind;

/* followRecoil() policies of the in-tree Trim subclasses. */
typedef enum {
  MTB_FOLLOW_ALL = 0,   /* TrimBase::followRecoil (trim.C:427-437), ThreadedTrimBase with primaries_only=false */
  MTB_FOLLOW_NONE = 1,  /* primaries_only=true; TrimRange (TrimRange.h:16) */
  MTB_FOLLOW_GEN_LT = 2 /* TrimPrimaries/TrimRecoils: recoil->_gen < follow_max_gen (trim.h:120,133) */
} mtb_follow;

/* vacancyCreation() variants. */
typedef enum {
  MTB_VAC_COUNT = 0, /* vacancies_created++ (trim.C:439-443; TrimVacCount.C:33; TrimVacEnergyCount.C:33) */
  MTB_VAC_NRT = 1,   /* TrimRange NRT damage estimate (apps/src/TrimRange.C:31-47) */
  MTB_VAC_KP = 2,    /* ++ plus modified Kinchin-Pease for gen == follow_max_gen (trim.C:445-464) */
  MTB_VAC_NONE = 3   /* subclasses that do not touch the counter (TrimVacMap, TrimDefectLog) */
} mtb_vacancy_model;

/* Tally bit mask: which hook side effects of the in-tree subclasses run on the device. */
enum {
  MTB_TALLY_VAC_DEPTH = 1u << 0,  /* TrimVacCount vac/repl depth histograms (TrimVacCount.C:31-53) */
  MTB_TALLY_VAC_ENERGY = 1u << 1, /* TrimVacEnergyCount [int ln E][int x] (TrimVacEnergyCount.C:31-53) */
  MTB_TALLY_RANGE = 1u << 2,      /* TrimRange::dissipateRecoilEnergy x-list per Z (TrimRange.C:49-54) */
  MTB_TALLY_PHONON = 1u << 3,     /* TrimPhononOut EnucTotal bookkeeping (trim.C:503-527) */
  MTB_TALLY_VACMAP = 1u << 4,     /* TrimVacMap 20x20x3 (trim.C:483-501) */
  MTB_TALLY_RECORDS = 1u << 5,    /* one mtb_record per primary (for statistics) */
  MTB_TALLY_IONLOG = 1u << 6      /* one mtb_ion_log per followed ion passing the Z filter */
};

/* Run-wide constants: the SimconfType fields trim() reads (simconf.h:66, simconf.C:44-49)
 * plus which hook behaviour to apply. */
typedef struct {
  double tmin;          /* 0.2 */
  double tau;           /* 0.0 */
  double cw;            /* 0.001 */
  double length_scale;  /* 1.0  (SimconfType::setLengthScale) */
  int32_t potential;    /* mtb_potential */
  int32_t follow;       /* mtb_follow */
  int32_t follow_max_gen;
  int32_t vacancy_model; /* mtb_vacancy_model */
  uint32_t tally_mask;
  int32_t vmap_z[3];    /* TrimVacMap z1,z2,z3 */
  int32_t ionlog_z;     /* log only ions with this Z; 0 = all */
  int32_t hist_bins;    /* depth bins kept for VAC_DEPTH / VAC_ENERGY; 0 = max(16384, 4*extent); deeper
                           events land in the last bin and are counted in mtb_counters.hist_clamped */
  int32_t evac_rows;    /* ln(E) rows kept for VAC_ENERGY; 0 = 32 */
  int32_t device;       /* CUDA device ordinal */
  uint64_t ionlog_capacity; /* entries; 0 = 1<<20 */
  uint64_t range_capacity;  /* entries; 0 = 1<<22 */
} mtb_config;

/* element.h:29-41 (input part). */
typedef struct {
  int32_t Z;
  int32_t _pad;
  double m;      /* amu */
  double t;      /* relative amount, normalised by prepare() (material.C:36-54) */
  double Edisp;  /* 25 eV default (element.C:25) */
  double Elbind; /* 3 eV default */
} mtb_element;

/* material.h:34-75 (input part).  Elements of material i are
 * elements[first_element .. first_element + n_elements). */
typedef struct {
  double rho;    /* g/cm^3 */
  int32_t tag;   /* MaterialBase::_tag, -1 default (material.C:31) */
  int32_t n_elements;
  int32_t first_element;
  int32_t _pad;
} mtb_material;

/* sample.h:33-58 + subclass parameters. */
typedef struct {
  int32_t kind;  /* mtb_geometry_kind */
  int32_t bc[3]; /* mtb_boundary per axis */
  double w[3];   /* simulation volume */
  /* LAYERS: material index == layer index (sample_layers.C:26-49) */
  int32_t n_layers;
  int32_t _pad0;
  const double * layer_thickness;
  /* CLUSTERS: matrix = material 0, inclusions = material 1 (sample_clusters.C:43-55) */
  int32_t kn[3]; /* spatial hash dimensions (initSpatialhash) */
  int32_t n_clusters;
  const double * cluster_xyzr; /* 4 doubles per cluster: x, y, z, r — in addCluster() order */
} mtb_geometry;

/* ion.h:45-84. */
typedef struct {
  double pos[3];
  double dir[3];
  double E;   /* eV */
  double m;   /* amu */
  double Ef;  /* 3 eV default (ion.C:26) */
  int32_t Z;
  int32_t gen;
  int32_t tag;
  uint32_t seed; /* IonBase::_seed; used by the mt19937 oracle mode only */
} mtb_ion;

/* Per-primary record (MTB_TALLY_RECORDS): everything the statistical criteria need. */
typedef struct {
  double pos[3];  /* final position of the primary ion itself */
  double E;       /* its final energy */
  double Eel;     /* electronic loss of the whole cascade */
  double Enuc;    /* phonon/binding losses (only with MTB_TALLY_PHONON) */
  uint32_t vacancies;
  uint32_t replacements;
  uint32_t steps; /* collision steps of the whole cascade */
  uint32_t ions;  /* ions followed, primary included */
  int32_t state;  /* mtb_ion_state of the primary when it stopped */
  uint32_t primary_steps;
} mtb_record;

/* Per-ion log entry (MTB_TALLY_IONLOG): birth and death of a followed ion. */
typedef struct {
  double pos0[3];
  double pos1[3];
  double E0;
  double E1;
  uint64_t uid;     /* scheduling-independent ion id (Philox stream id) */
  uint64_t primary; /* global primary index */
  int32_t Z;
  int32_t gen;
  int32_t tag;
  int32_t state;
} mtb_ion_log;

/* simconf.h:67,104-109 plus bookkeeping the reference does not keep. */
typedef struct {
  uint64_t vacancies_created;
  uint64_t replacements;
  uint64_t steps;        /* collision steps (iterations of trim.C:74-424) */
  uint64_t ions;         /* trim() calls */
  uint64_t primaries;
  uint64_t recoils_queued;
  uint64_t lost;         /* ions that crossed a CUT boundary */
  uint64_t left_sample;  /* ions that flew into vacuum (lookupMaterial == 0) */
  uint64_t hist_clamped; /* depth tallies beyond hist_bins (counted in the last bin) */
  uint64_t stack_max;    /* deepest per-lane recoil stack seen */
  double EelTotal;
  double EnucTotal;
} mtb_counters;

/* One collision of a single ion, as seen by the five virtual hooks (trim.C:357-418).
 * Lets a host replay arbitrary TrimBase subclasses. */
typedef struct {
  double pka_pos[3];    /* after the free flight */
  double pka_dir[3];
  double pka_E;
  double recoil_pos[3]; /* previous collision site (trim.C:306-310) */
  double recoil_dir[3]; /* normalised if recoil_above_threshold */
  double recoil_E;      /* den - Elbind */
  double ls, dee, den;
  int32_t material;     /* index into materials */
  int32_t element;      /* index into that material's elements */
  int32_t material_tag;
  int32_t pka_state;    /* state after the fate decision, before followRecoil veto */
  int32_t recoil_above_threshold; /* recoil.E > Edisp - Elbind */
  int32_t _pad;
} mtb_event;

typedef struct mtb_handle mtb_handle;

/* Library / device introspection. */
const char * mtb_version(void);
const char * mtb_last_error(void);
int mtb_device_count(void);

void mtb_default_config(mtb_config * cfg);

/* Replaces `new SimconfType` + `new Trim*`: creates an engine bound to one GPU. */
int mtb_create(const mtb_config * cfg, mtb_handle ** out);
int mtb_destroy(mtb_handle * h);

/* Replaces SimconfType::readDataFiles (simconf.C:81-137) for the 11 hot-path columns.
 * Optional: the ZBL-85 tables are built in.  Arrays are indexed by Z-1. */
int mtb_set_tables(mtb_handle * h, const double * pcoef /*[92][8]*/, const double * vfermi /*[92]*/,
                   const double * lfctr /*[92]*/, const double * mm1 /*[92]*/);
int mtb_get_tables(double * pcoef, double * vfermi, double * lfctr, double * mm1);

/* Replaces MaterialBase construction + prepare() (material.C:36-74). */
int mtb_set_materials(mtb_handle * h, int n_materials, const mtb_material * materials,
                      int n_elements, const mtb_element * elements);
/* Replaces the SampleBase subclass object and its lookupMaterial(). */
int mtb_set_geometry(mtb_handle * h, const mtb_geometry * geom);

/* The hot path.  Replaces, for n primaries, the loop
 *   queue.push(pka); while(!queue.empty()){ r=pop; sample->averages(r); trim->trim(r,queue); }
 * (apps/runmytrim.C:76-92).  Primary i uses Philox stream id first_index+i under key `seed`,
 * so results do not depend on how primaries are sharded.  Tallies accumulate in the handle.
 * `records` (optional, n entries) needs MTB_TALLY_RECORDS. */
int mtb_run(mtb_handle * h, uint64_t n, const mtb_ion * primaries, uint64_t seed,
            uint64_t first_index, mtb_record * records);
/* Same, for n copies of one template ion (no per-primary host array). */
int mtb_run_beam(mtb_handle * h, uint64_t n, const mtb_ion * ion, uint64_t seed,
                 uint64_t first_index, mtb_record * records);

/* Split form for callers that keep primaries resident in HBM: upload once, launch many. */
int mtb_upload_primaries(mtb_handle * h, uint64_t n, const mtb_ion * primaries);
int mtb_launch_resident(mtb_handle * h, uint64_t seed, uint64_t first_index); /* asynchronous */
int mtb_synchronize(mtb_handle * h);
/* Device time of the most recent transport kernel launch(es) of mtb_run/mtb_launch_resident, in ms. */
int mtb_last_kernel_ms(mtb_handle * h, float * ms);
/* Name of the compile-time kernel variant the handle's configuration selects: "MONO" (a launch
 * without per-primary records runs its "MONO-NOREC" twin), "FAST", "MONO-EVAC", "FAST-PHONON", "CLUSTERS-LOG",
 * "CLUSTERS", "LAYERS-PLAIN", "LAYERS" or "GENERIC"; "" on error.  Diagnostic: the reference has no counterpart. */
const char * mtb_kernel_variant(mtb_handle * h);
int mtb_fetch_records(mtb_handle * h, uint64_t n, mtb_record * records);

/* Tally read-back (replaces threadJoin + the accessors writeOutput uses). */
int mtb_reset_tallies(mtb_handle * h);
int mtb_get_counters(mtb_handle * h, mtb_counters * out);
/* Depth histograms; *n_bins receives the number of bins up to the last non-zero one. */
int mtb_get_vac_depth(mtb_handle * h, uint64_t * vac, uint64_t * repl, size_t capacity, size_t * n_bins);
int mtb_get_vac_energy(mtb_handle * h, uint64_t * evac /*[rows][bins]*/, size_t rows, size_t bins);
int mtb_get_vacmap(mtb_handle * h, uint64_t * vmap /*[20][20][3]*/);
int mtb_get_range_list(mtb_handle * h, float * x, int32_t * Z, size_t capacity, size_t * n);
int mtb_get_ion_log(mtb_handle * h, mtb_ion_log * out, size_t capacity, size_t * n);
/* Empties the ion log and the range list (the other tallies keep accumulating). */
int mtb_clear_lists(mtb_handle * h);
int mtb_hist_bins(mtb_handle * h, size_t * bins, size_t * evac_rows);

/* Raw device views of the additive tallies so a caller can reduce them across GPUs with
 * its own collective (torch.distributed / NCCL): u64 block and f64 block. */
int mtb_tally_device_views(mtb_handle * h, void ** u64_dev, size_t * n_u64, void ** f64_dev, size_t * n_f64);
/* Single-process multi-GPU join of the additive tallies (replaces ThreadedTrimBase::threadJoin, runmytrim.C:316-323):
 * NCCL all-reduce over the handles' tally blocks, IN PLACE on every handle — afterwards each handle holds the job
 * totals (sum of the counters, histograms and energies; max of the stack high-water mark).  Call it once, when all
 * handles have finished: a second call, or more primaries followed by another call, would add the totals up again. */
int mtb_allreduce(mtb_handle ** handles, int n_handles);

/* Fission-fragment source of the UO2 experiment (replaces the event loop head of apps/mytrim_uo2.C:226-270): events
 * [first_event, first_event + n_events) of the sequence a std::mt19937 seeded with `seed` produces — fragment mass from
 * the cumulative fission yield (MassInverter, invert.C:30-56), total kinetic energy (EnergyInverter, invert.C:59-78),
 * Z1 = round(92 A1 / 235), isotropic back-to-back directions, uniform origin in the box w.  Writes 2 n_events
 * primaries (gen 0, tag -1, Ef = 3 eV as IonBase::setEf gives them) and returns their summed energy in *e_total.
 * About one draw in 1e6 ends the reference's 32-step bisection at A1 = 235 * 2^-33, i.e. Z1 = 0, for which the
 * reference's stopping reads scoef[-1] (undefined behaviour): such a fragment keeps its place in the list (its index is
 * its Philox stream) with Z = 1, m = 1 and NO energy — it stops where it starts — and a note goes to stderr.
 * Host only; sharding a run = giving every GPU its own event range. */
int mtb_fission_pairs(uint32_t seed, uint64_t first_event, uint64_t n_events, const double w[3], mtb_ion * out,
                      double * e_total);

/* Measures the FP32 FMA issue rate of `device` (the roofline denominator of this path). */
int mtb_measure_fp32_peak(int device, double * tflops, float * ms);

/* Replaces one TrimBase::trim(pka, recoils) call for arbitrary subclasses: follows ONE ion,
 * never follows recoils, and reports every collision so the host can run the virtual hooks
 * and fill its own std::queue.  `ion` is updated in place (final pos/dir/E/state). */
int mtb_trim_one(mtb_handle * h, mtb_ion * ion, uint64_t seed, uint64_t uid, int32_t * final_state,
                 mtb_event * events, size_t capacity, size_t * n_events);

/* The same for a batch of ions in ONE launch (one GPU lane per ion; replaces the per-ion round trip when a reference
 * app's queue loop, runmytrim.C:76-92, hands TrimBase::trim() one ion at a time: the façade follows all queued ions of
 * a generation at once).  Ion i uses Philox stream uids[i] (first_uid + i when uids is null) and writes its collisions to
 * events[i * events_per_ion ...]; counts[i] is the number of collisions it had.  Where counts[i] > events_per_ion the
 * record of that ion is incomplete (its first events_per_ion events are valid): follow it again with the same stream
 * id — mtb_trim_one or another batch with that ion's uid — and a larger buffer; the replay is identical.
 * ions[i] and final_states[i] receive the final state of every completely recorded ion. */
int mtb_trim_many(mtb_handle * h, size_t n, mtb_ion * ions, uint64_t seed, uint64_t first_uid, const uint64_t * uids,
                  int32_t * final_states, mtb_event * events, size_t events_per_ion, uint32_t * counts);

/* Replaces MaterialBase::getrstop (material.C:113-122) for a batch of (Z1, m1, E) in material
 * `material`; runs the same device function the transport kernel uses. */
int mtb_stopping(mtb_handle * h, int material, size_t n, const int32_t * Z1, const double * m1,
                 const double * E, double * out);

#ifdef __cplusplus
}
#endif
#endif /* MYTRIM_B200_H */
