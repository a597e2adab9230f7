// mtb_types.h — flattened, device-resident form of the MyTRIM plugin objects.
//
// The host side (mtb_capi.cu) turns {SimconfType tables, MaterialBase/Element vectors, a
// SampleBase subclass + parameters, the Trim subclass' hook behaviour} into these PODs once per
// configuration; kernels copy the small tables into shared memory at start-up.
#ifndef MTB_TYPES_H
#define MTB_TYPES_H

#include <stdint.h>
#include "../../include/mytrim_b200.h"

namespace mtb
{

// One target element of one material.  Everything MaterialBase::average (material.C:99-108) and
// MaterialBase::rstop/rpstop (material.C:133-282) need about the TARGET atom, in float.
struct DevElement
{
  float m;      // amu
  float t;      // normalised stoichiometric fraction
  float Edisp;  // eV
  float Elbind; // eV
  float z023;   // Z^0.23 (universal screening length)
  float fz;     // float(Z)
  float vfermi; // scoef[Z-1].vfermi
  float vf2inv; // 1/(2 vfermi^2)
  float pc[8];  // scoef[Z-1].pcoef[0..7]
  int32_t Z;
  float velpwr; // 0.25 (Z<=6) or 0.45: rpstop low-energy exponent (material.C:151-155)
  int32_t zslot; // column of this element's Z in the low-velocity stopping table
  int32_t tcls;  // target class: index of this element's (Z, m) among the distinct target atoms
  float sp25;    // proton stopping at 25 keV/amu, the value rpstop scales below 25 keV/amu (material.C:137-157)
  float _pad[3];
};
static_assert(sizeof(DevElement) == 96, "DevElement layout");

// One material (prepare() results, material.C:36-74).
struct DevMaterial
{
  float arho;  // atoms/Ang^3
  float am;    // mean mass
  float az;    // mean Z
  float az023; // az^0.23
  int32_t n_elem;
  int32_t first_elem;
  int32_t tag;
  int32_t user_index; // index in the caller's material list (before de-duplication)
};
static_assert(sizeof(DevMaterial) == 32, "DevMaterial layout");

// Low-velocity heavy-ion stopping of projectile Z1 in target Z2 (one entry per (Z1, distinct Z2)).
// In the velocity-proportional regime of MaterialBase::rstop (material.C:259-273, yr clamped at
// yrmin) everything except (e/eee)^power depends on (Z1, Z2) only:
//     rstop = coef * e^power        for Z1 >= 3 and e = E/(1000 m1) <= e_max  [keV/amu]
// coef is evaluated on the host in double precision.
struct LowStop
{
  float coef;  // 10 * rpstop(Z2, eee) * (zeta * Z1)^2 / eee^power
  float e_max; // min(e at which yr leaves the clamp, 20 keV/amu); 0 disables the shortcut
  float power; // 0.5 or 0.375
  float pad;
};

// Projectile class: a distinct (Z, m) that can be in flight — every target class (recoils) plus the
// species of the primaries the host has seen.  Tabulating per-class constants on the host (double
// precision, rounded once) replaces MaterialBase::average (material.C:77-110) on the device.
struct ProjClass
{
  float m2;     // 2 m  (momentum = sqrt(m2 * E))
  float inv_km; // 0.001 / m: eV -> keV/amu
  float m;
  float fz;     // float(Z)
  float z023;   // Z^0.23
  float cbrt;   // Z^(1/3)
  float lfctr;  // screening length factor of Z
  int32_t Z;
};
static_assert(sizeof(ProjClass) == 32, "ProjClass layout");

// (projectile class, material): free-flight constants (material.C:80-93, trim.C:88-92)
//   eeg = K sqrt(E);  D = eeg + sqrt(eeg) + 0.125 eeg^0.1;  pmax = a / D;  ls = C2 D^2
struct PairM
{
  float a;  // screening length
  float K;  // sqrt(f * epsdg)
  float C2; // 1 / (pi arho a^2)
  float sk; // sqrt(0.001 / m1): sqrt(E) * sk = sqrt(e), e in keV/amu (shares the sqrt(E) of the free flight)
};

// (projectile class, target class): collision constants (material.C:99-108)
struct PairE
{
  float my;     // m1 / m2
  float ec;     // 4 my / (1 + my)^2
  float inv_ai; // 1 / screening length
  float sfi;    // sqrt of the reduced-energy factor fi: sqrt(eps) = sfi * sqrt(E)
};

// Per projectile-Z constants (indexed by Z, entry 0 unused).
struct DevIonZ
{
  float z023;  // Z^0.23
  float cbrt;  // Z^(1/3)
  float lfctr; // scoef[Z-1].lfctr
  float mm1;   // scoef[Z-1].mm1 (used when an ion has mass 0, material.C:181-184)
};

// Suspended ion on a lane's private stack: exactly two 32-byte sectors.
struct __attribute__((aligned(16))) StackEntry
{
  double pos[3];
  double E;
  float dir[3];
  uint32_t ic;     // collision steps already taken by this ion
  uint64_t uid;    // Philox stream id
  uint32_t packed; // species (0: the primary, else 1 + projectile class) | gen << 12 | flags << 28
  int32_t tag;
};
static_assert(sizeof(StackEntry) == 64, "StackEntry layout");

enum
{
  SPECIES_PRIMARY = 0,   // (Z, m) of the lane's current primary
  SPECIES_CLASS0 = 1,    // 1 + projectile class of the target atom the recoil was (DevElement::tcls)
  SPECIES_MASK = 0xFFF,
  GEN_SHIFT = 12,
  GEN_MASK = 0xFFFF,
  FLAG_PRIMARY = 1u << 28,
  // bits 29..31 of a SUSPENDED ion (stack / sharing pool entry): the path it may still fly before the next cluster
  // look-up can matter, in whole units of LaunchParams::cl_safe_unit, saturated at 7 (clusters geometry only)
  SAFE_SHIFT = 29,
  SAFE_MAX = 7
};

#define MTB_STACK_DEPTH 32

// Indices into the u64 counter block (mtb_counters order, then histograms).
enum
{
  CNT_VAC = 0,
  CNT_REPL,
  CNT_STEPS,
  CNT_IONS,
  CNT_PRIMARIES,
  CNT_QUEUED,
  CNT_LOST,
  CNT_LEFT,
  CNT_CLAMPED,
  CNT_STACKMAX,
  CNT_NEXT_PRIMARY, // work counter (not reduced)
  CNT_IONLOG_N,
  CNT_RANGE_N,
  CNT_ERROR,
  CNT_EVENTS_N,
  CNT_DEFERRED, // primaries the fast kernel handed to the generic kernel (per-primary species)
  CNT_COUNT = 16
};

// 16-byte row of the class tables (ProjClass = 2 rows, PairM = 1, PairE = 1)
struct __attribute__((aligned(16))) float4_t
{
  float x, y, z, w;
};

// Slot of a CTA's work-sharing pool in shared memory (bounded multi-producer multi-consumer ring,
// D. Vyukov's scheme): a suspended ion plus the primary it belongs to.
struct __attribute__((aligned(16))) PoolSlot
{
  unsigned long long seq;
  uint64_t prim;
  StackEntry e;
};
static_assert(sizeof(PoolSlot) == 80, "PoolSlot layout");

enum
{
  POOL_ENQ = 0,     // enqueue ticket
  POOL_DEQ = 1,     // dequeue ticket
  POOL_WORKING = 2, // lanes that hold work + entries in the pool; 0 <=> nothing left anywhere
  POOL_IDLE = 3,    // lanes polling the pool
  POOL_CTL_COUNT = 4
};
#ifndef MTB_POOL_SLOTS
#define MTB_POOL_SLOTS 32 // per CTA, power of two
#endif

struct RangeEntry
{
  float x;
  int32_t Z;
};

// Everything a transport launch needs, passed by value (lives in the constant bank).
struct LaunchParams
{
  // run constants
  float tmin, tau, cw, inv_scale;
  int32_t potential, follow, follow_max_gen, vacancy_model;
  uint32_t tally_mask;
  int32_t vmap_z[3];
  int32_t ionlog_z;
  // tables (global memory, copied to shared at kernel start)
  const DevElement * elements;
  const DevMaterial * materials;
  const DevIonZ * ionz; // [93]
  const LowStop * lowstop; // [93][n_zslots]
  int32_t n_zslots;
  const ProjClass * pclass; // [n_pclass]: target classes first, then primary species
  const PairM * pairm;      // [n_pclass][n_materials]
  const PairE * paire;      // [n_pclass][n_tclass]
  int32_t n_pclass, n_tclass;
  const int32_t * tclass_elem; // [n_tclass]: one element index per target class
  float4_t * custom_rows;      // [lanes][2 + n_materials + n_tclass]: class rows of per-primary species
  int32_t n_elements, n_materials;
  // geometry
  int32_t geom_kind;
  int32_t bc[3];
  double w[3];
  int32_t n_layers;
  const double * layer_cum;     // cumulative thickness, [n_layers]
  const int32_t * layer_mat;    // de-duplicated material id per layer
  int32_t kn[3];
  int32_t cl_ks[3];             // int(cmr/kd)+1 per axis
  const int32_t * cl_hash;      // sh[]
  const int32_t * cl_next;      // cl[]
  const double * cl_xyzr;       // 4 per cluster
  double kd[3];
  double inv_kn[3];             // 1 / kn
  double kn_w[3];               // kn / w: cell index without a division (exact path on near-ties)
  const uint8_t * cl_dist;      // per hash cell: chessboard distance (in cells, capped at 255) to the nearest cell
                                // whose scan neighbourhood holds a cluster; 0 = scan here
  const float * cl_safe;        // per hash cell: lower bound of the distance from any point of the cell to the nearest
                                // cluster surface (fully periodic boxes; else null), mtb_tables.h
  float cl_inv_safe_unit;       // 1 / cl_safe_unit, 0 when cl_safe_unit is 0
  float cl_safe_unit;           // path length an ion can travel per unit of cl_dist above 1 without meeting a
                                // cluster (smallest cell edge, with a margin); 0 = no skipping (non-periodic box)
  // primaries
  const mtb_ion * primaries;    // device copy, or null for beam mode
  mtb_ion beam;
  uint64_t n_primaries, first_index;
  const uint32_t * index_list;  // optional: launch over primaries[index_list[k]], k < n_primaries
  uint32_t * deferred;          // fast kernel: indices of primaries without a projectile class
  uint32_t key0, key1;
  float share_min_E; // work sharing: smallest energy [eV] of a pair of ions one of which may be donated
  uint32_t rk[20]; // Philox round keys (philox_round_keys): the key schedule is a launch constant
  // outputs
  unsigned long long * u64;     // counter + histogram block
  double * f64;                 // [0]=Eel, [1]=Enuc
  int32_t hist_bins, evac_rows;
  int32_t smem_hist_bins;       // depth bins mirrored in shared memory
  int32_t mono;                 // one material made of one element in a solid/layered sample (host flag: MONO variants)
  mtb_record * records;         // [n_primaries] or null
  mtb_ion_log * ionlog;
  unsigned long long ionlog_cap;
  RangeEntry * range;
  unsigned long long range_cap;
  StackEntry * stacks;          // [lanes][MTB_STACK_DEPTH]
  // single-ion event mode
  mtb_event * events;            // event mode: lane i writes events[i * events_cap ...]
  unsigned long long events_cap; // per ion
  uint64_t single_uid;           // event mode: stream id of ion i = single_uid + i
  const uint64_t * uid_list;     // event mode: explicit stream ids (else single_uid + i)
  uint32_t * event_counts;       // event mode with several ions (mtb_trim_many): collisions of ion i (may exceed events_cap)
  int32_t one_material;         // solid/layered sample whose layers are all the same material (host flag: no layer search)
  int32_t n_input_materials;    // materials the caller passed (identical ones are folded on the device, mtb_tables.h)
};

// offsets inside the u64 block
inline size_t
u64_off_vac(const LaunchParams &)
{
  return CNT_COUNT;
}

} // namespace mtb
#endif
