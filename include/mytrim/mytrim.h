// mytrim.h — C++ plugin surface of MyTRIM, re-created over the B200 engine.
//
// Source-compatible with the reference's public headers (simconf.h, ion.h, element.h,
// material.h, sample*.h, trim.h, invert.h, functions.h and the apps' ThreadedTrimBase /
// TrimVacCount / TrimVacEnergyCount / TrimRange): same namespace, class names, public data
// members and virtual hooks, so an app such as runmytrim or mytrim_layers recompiles against it
// (SURVEY.md §8b).  What changed is where the work happens:
//
//   * TrimBase::trim(pka, recoils) — one ion — runs on the GPU through mtb_trim_one(); the five
//     virtual hooks are then replayed on the host from the per-collision event records, in the
//     reference's order, so arbitrary user subclasses keep working.
//   * TrimBase::trimBatch(primaries) — new — hands whole cascades (the app's
//     pop/averages/trim/delete loop, runmytrim.C:76-92) to the persistent transport kernel and
//     keeps the tallies of the in-tree subclasses on the device.  Subclasses describe their hook
//     behaviour through deviceHooks(); the default says "unknown subclass" and trimBatch() refuses
//     to run rather than silently computing something else.
//   * Samples describe their geometry through SampleBase::describe(); user-defined
//     lookupMaterial() overrides cannot run on the device and are rejected the same way.
//
// There is no CPU transport path in this library.
#ifndef MYTRIM_B200_FACADE_H
#define MYTRIM_B200_FACADE_H

#include <cmath>
#include <fstream>
#include <iostream>
#include <memory>
#include <queue>
#include <random>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../mytrim_b200.h"

typedef double Real;

// ---- shim/point.h -------------------------------------------------------------------------
class Point
{
public:
  Point() : _v{0.0, 0.0, 0.0} {}
  Point(Real x, Real y, Real z) : _v{x, y, z} {}
  Real & operator()(unsigned int i) { return _v[i]; }
  const Real & operator()(unsigned int i) const { return _v[i]; }
  Real norm_sq() const { return _v[0] * _v[0] + _v[1] * _v[1] + _v[2] * _v[2]; }
  Real norm() const { return std::sqrt(norm_sq()); }
  Point operator+(const Point & o) const { return Point(_v[0] + o._v[0], _v[1] + o._v[1], _v[2] + o._v[2]); }
  Point operator-(const Point & o) const { return Point(_v[0] - o._v[0], _v[1] - o._v[1], _v[2] - o._v[2]); }
  Point operator*(Real s) const { return Point(_v[0] * s, _v[1] * s, _v[2] * s); }
  Point operator/(Real s) const { return Point(_v[0] / s, _v[1] / s, _v[2] / s); }
  Point operator-() const { return Point(-_v[0], -_v[1], -_v[2]); }
  Point & operator+=(const Point & o) { for (int i = 0; i < 3; ++i) _v[i] += o._v[i]; return *this; }
  Point & operator-=(const Point & o) { for (int i = 0; i < 3; ++i) _v[i] -= o._v[i]; return *this; }
  Point & operator*=(Real s) { for (int i = 0; i < 3; ++i) _v[i] *= s; return *this; }
  Point & operator/=(Real s) { for (int i = 0; i < 3; ++i) _v[i] /= s; return *this; }

private:
  Real _v[3];
};

// ---- shim/pow.h ---------------------------------------------------------------------------
namespace Utility
{
template <int N, typename T>
inline T
pow(const T & x)
{
  T r = 1, b = x;
  for (int n = N; n > 0; n >>= 1)
  {
    if (n & 1)
      r = r * b;
    b = b * b;
  }
  return r;
}
} // namespace Utility

namespace MyTRIM_NS
{

/// Thrown by the façade where the reference signature leaves no room for a status (TrimBase::trim(),
/// MaterialBase::getrstop(), SimconfType's table reader): no GPU, a sample or hook set the device cannot run, a CUDA
/// error.  what() carries mtb_last_error().  Nothing in the library calls exit().
class EngineError : public std::runtime_error
{
public:
  explicit EngineError(const std::string & what) : std::runtime_error(what) {}
};


// ---- functions.h --------------------------------------------------------------------------
inline void v_cross(const Real * a, const Real * b, Real * c)
{
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
inline void v_cross(const Point & a, const Point & b, Point & c)
{
  c = Point(a(1) * b(2) - a(2) * b(1), a(2) * b(0) - a(0) * b(2), a(0) * b(1) - a(1) * b(0));
}
inline void v_scale(Real * a, Real s) { a[0] *= s; a[1] *= s; a[2] *= s; }
inline Real v_dot(const Real * a, const Real * b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline void v_norm(Real * a, Real len = 1.0) { v_scale(a, len / std::sqrt(v_dot(a, a))); }
inline void v_norm(Point & a, Real len = 1.0) { a *= len / a.norm(); }
inline Real sqr(Real a) { return a * a; }
inline Real cub(Real a) { return a * a * a; }

class IonBase;
class MaterialBase;
class SampleBase;
class TrimBase;

// ---- simconf.h ----------------------------------------------------------------------------
class SimconfType
{
public:
  SimconfType(unsigned int seed = 12345678);
  ~SimconfType();

  Real drand() { return _uniform(*_rng); }
  unsigned int irand() { return _uniform_int(*_rng); }
  void seed(unsigned int seed);

  void setLengthScale(Real l);
  Real lengthScale() { return _length_scale; }
  Real areaScale() { return _area_scale; }
  Real volumeScale() { return _volume_scale; }

  Real ed, tmin, tau, da, cw;
  int _id;

  struct ScoefLine
  {
    ScoefLine();
    std::string sym, name;
    Real mm1, m1, mnat, rho, atrho, vfermi, heat, lfctr;
    std::vector<Real> pcoef, ehigh, screen, fermicorr;
  };
  static const unsigned int _rows = 92;
  std::vector<ScoefLine> scoef;
  ScoefLine scoeflast;
  Real snuc[_rows][_rows][4];

  bool fullTraj;
  Real EelTotal;
  Real EnucTotal;
  int vacancies_created;

  // ---- B200 engine plumbing (not in the reference) ----
  /// 64-bit key of the Philox streams; seed() sets it to the seed given.
  uint64_t philoxKey() const { return _philox_key; }
  /// next unused stream id; every trim()/trimBatch() primary consumes one.
  uint64_t nextStreamId(uint64_t n = 1) { const uint64_t v = _stream; _stream += n; return v; }
  /// A driver that shards one run over several SimconfType objects (one per GPU) gives every shard the global
  /// index of its first primary: results then do not depend on the split.
  void setStreamId(uint64_t first) { _stream = first; }
  /// CUDA device the engines created from this SimconfType run on.
  int device = 0;

private:
  void loadTables();
  std::string _data_dir;
  std::unique_ptr<std::mt19937> _rng;
  std::uniform_real_distribution<double> _uniform;
  std::uniform_int_distribution<unsigned int> _uniform_int;
  Real _length_scale, _area_scale, _volume_scale;
  uint64_t _philox_key, _stream;
};

extern SimconfType * simconf;

// ---- ion.h --------------------------------------------------------------------------------
class IonBase
{
public:
  IonBase();
  IonBase(IonBase * prototype);
  IonBase(int Z, Real m, Real E);
  virtual ~IonBase() {}

  virtual void parent(IonBase * parent);
  virtual IonBase * spawnRecoil();
  void setEf();
  bool operator<(const IonBase &) const;

  int _Z;
  Real _m;
  Real _E;
  Point _dir, _pos;
  unsigned int _seed;
  int _gen, _id;
  int _tag;
  Real _Ef;
  enum StateType { MOVING, REPLACEMENT, SUBSTITUTIONAL, INTERSTITIAL, LOST, DELETE, VACANCY } _state;
};
std::ostream & operator<<(std::ostream & os, const IonBase & i);

class IonMDTag : public IonBase
{
public:
  IonMDTag() : IonBase(), _md(0) {}
  IonMDTag(IonMDTag * prototype) : IonBase(prototype), _md(prototype->_md) {}
  virtual IonBase * spawnRecoil();
  int _md;
};
std::ostream & operator<<(std::ostream & os, const IonMDTag & p);

class IonClock : public IonBase
{
public:
  IonClock() : IonBase(), _time(0.0) {}
  IonClock(IonClock * prototype) : IonBase(prototype), _time(prototype->_time) {}
  void parent(IonBase * parent);
  Real _time;
};

// ---- element.h ----------------------------------------------------------------------------
class Element
{
public:
  Element();
  int _Z;
  Real _m, _t;
  Real _Edisp, _Elbind;
  Real my, ec, ai, fi;
};

// ---- material.h ---------------------------------------------------------------------------
class MaterialBase
{
public:
  MaterialBase(SimconfType * simconf, Real rho);
  virtual ~MaterialBase();

  void prepare();
  void average(const IonBase * pka);
  /// electronic stopping [eV/Ang]; evaluated by the device function the transport kernel uses
  Real getrstop(const IonBase * pka);
  Real getDrstopDcomp(const IonBase * pka, const Element & component);
  virtual const Element & getElement(unsigned int nn) { return _element[nn]; }

  Real _rho;
  Real _am, _az;
  Real _arho;
  Real mu;
  Real a, f, epsdg;
  Real fd, kd;
  Real pmax;
  int _tag;
  bool _dirty;
  std::vector<Element> _element;

protected:
  SimconfType * _simconf;

private:
  mtb_handle * _engine; // lazily created single-material engine behind getrstop()
  std::vector<Element> _engine_elements;
  Real _engine_rho;
};

// ---- sample.h -----------------------------------------------------------------------------
class SampleBase
{
public:
  SampleBase(Real x = 10000.0, Real y = 10000.0, Real z = 10000.0);
  virtual ~SampleBase() {}

  virtual void averages(const IonBase * pka);
  virtual MaterialBase * lookupMaterial(Point & pos) = 0;
  virtual Real rangeMaterial(Point & pos, Point & dir);

  /// Geometry description for the device lookup.  Returns false for sample types the device
  /// does not know (user subclasses with their own lookupMaterial): such samples cannot be used
  /// with the B200 engine and TrimBase reports an error instead of falling back to the CPU.
  /// `storage` keeps arrays referenced by `out` alive.
  virtual bool describe(mtb_geometry & out, std::vector<double> & storage) const;

  std::vector<MaterialBase *> material;
  Real w[3];
  enum sampleBoundary { PBC, INF, CUT };
  sampleBoundary bc[3];

protected:
  void describeBox(mtb_geometry & out, int kind) const;
};

class SampleSolid : public SampleBase
{
public:
  SampleSolid(Real x, Real y, Real z) : SampleBase(x, y, z) {}
  virtual MaterialBase * lookupMaterial(Point & pos);
  virtual bool describe(mtb_geometry & out, std::vector<double> & storage) const;
};

class SampleLayers : public SampleBase
{
public:
  SampleLayers(Real x, Real y, Real z) : SampleBase(x, y, z) {}
  virtual MaterialBase * lookupMaterial(Point & pos);
  virtual Real rangeMaterial(Point & pos, Point & dir);
  virtual int lookupLayer(Point & pos);
  virtual bool describe(mtb_geometry & out, std::vector<double> & storage) const;
  std::vector<Real> layerThickness;
};

class SampleWire : public SampleBase
{
public:
  SampleWire(Real x, Real y, Real z);
  virtual MaterialBase * lookupMaterial(Point & pos);
  virtual bool describe(mtb_geometry & out, std::vector<double> & storage) const;
};

class SampleBurriedWire : public SampleWire
{
public:
  SampleBurriedWire(Real x, Real y, Real z);
  virtual MaterialBase * lookupMaterial(Point & pos);
  virtual bool describe(mtb_geometry & out, std::vector<double> & storage) const;
};

struct sampleClusters : SampleBase
{
  Real sd, kd[3];
  int *sh, kn[3];
  int *cl, cn, cnm;
  Real * c[4];
  Real cmr;

  sampleClusters(Real x = 10000.0, Real y = 10000.0, Real z = 10000.0);
  ~sampleClusters();
  virtual MaterialBase * lookupMaterial(Point & pos);
  virtual bool describe(mtb_geometry & out, std::vector<double> & storage) const;

  int lookupCluster(Point & pos, Real dr = 0.0);
  void initSpatialhash(int x, int y, int z);
  void addCluster(Real x, Real y, Real z, Real r);
  void addRandomClusters(unsigned int n, Real r, Real dr, SimconfType * simconf);

protected:
  void clearClusters();
  void clearSpatialHash();
  void reallocClusters(int n);
};

// ---- trim.h -------------------------------------------------------------------------------
/// How a Trim subclass' hooks map onto device tallies (filled by TrimBase::deviceHooks()).
struct DeviceHooks
{
  bool known = false;        ///< false: hooks are arbitrary host code, only trim() (event replay) works
  int follow = MTB_FOLLOW_ALL;
  int follow_max_gen = 1;
  int vacancy_model = MTB_VAC_COUNT;
  unsigned tally_mask = 0;
  int vmap_z[3] = {-1, -1, -1};
  int ionlog_z = 0;
  int hist_bins = 0;
  unsigned long long ionlog_capacity = 0; ///< entries (birth + death halves); 0 = engine default
  unsigned long long range_capacity = 0;  ///< entries of the range list; 0 = engine default
  /// trimBatch() hands the primaries to the device in chunks of this many (0 = all at once) and calls
  /// collectDeviceTallies() after each: list-type tallies (range list) stay bounded however large the run is.
  unsigned long long batch_chunk = 0;
};

class TrimBase
{
public:
  TrimBase(SimconfType * simconf, SampleBase * sample);
  virtual ~TrimBase();

  /// One ion on the GPU; recoils the hooks decide to follow are pushed onto `recoils`.
  void trim(IonBase * pka, std::queue<IonBase *> & recoils);

  /// Whole cascades of all `primaries` on the GPU (replaces the app's queue loop).  Final
  /// position/energy/state are written back into the primaries.  Tallies accumulate on the
  /// device; SimconfType counters are updated; histograms are fetched by the subclasses'
  /// writeOutput()/accessors.  Returns false and sets lastError() if the subclass or the
  /// sample cannot run on the device.
  bool trimBatch(std::vector<IonBase *> & primaries);
  bool trimBatch(std::vector<IonBase *> & primaries, std::vector<mtb_record> * records);
  const std::string & lastError() const { return _error; }

  void setBaseName(const std::string & name) { _base_name = name; }
  virtual void writeOutput() {}

  enum Potential { UNIVERSAL, MOLIERE, CKR };
  Potential _potential;

  /// the engine handle (created on first use); exposed for multi-GPU reductions
  mtb_handle * engine();

protected:
  virtual bool followRecoil();
  virtual void vacancyCreation();
  virtual void replacementCollision() {}
  virtual void checkPKAState() {}
  virtual void dissipateRecoilEnergy() {}

  /// Describes the hooks above for the device.  Subclasses that override hooks must override
  /// this too, or leave `known` false.
  virtual void deviceHooks(DeviceHooks & h) const;
  /// Called by trimBatch() after the device tallies have been updated.
  virtual void collectDeviceTallies() {}
  /// Called when the batch engine was (re)created: its tallies restart at zero, so subclasses forget what they
  /// had already merged from the previous engine.
  virtual void resetDeviceBaselines() {}

  SimconfType * _simconf;
  SampleBase * _sample;
  IonBase *_pka, *_recoil;
  MaterialBase * _material;
  const Element * _element;
  std::queue<IonBase *> * recoil_queue_ptr;
  bool terminate;
  Real _ls;
  Real _dee;
  Real _den;
  std::string _base_name;

  bool ensureEngine(bool batch);
  std::string _error;

public:
  /// The engines snapshot SimconfType (tmin, tau, cw, length scale), _potential, the sample's materials and
  /// geometry and the hook description.  Every trim()/trimBatch() call compares a fingerprint of those with the
  /// snapshot and rebuilds the engine when something changed (the reference reads them on every call);
  /// invalidateEngine() forces the rebuild.
  void invalidateEngine();

private:
  // two engines: the batch engine keeps the device tallies of trimBatch() while trim() calls (single-ion event
  // mode, no tallies) come and go on their own handle
  mtb_handle * _engine;
  mtb_handle * _engine_single;
  unsigned long long _engine_fp, _engine_single_fp;
  unsigned long long configFingerprint(bool batch, const mtb_config & cfg, const std::vector<mtb_material> & mats,
                                       const std::vector<mtb_element> & els, const mtb_geometry & g,
                                       const std::vector<double> & storage) const;
  std::vector<mtb_event> _events;
  // trim() is handed one ion at a time, but the caller's queue shows which ions come next: every ion waiting there is
  // followed in the same launch (mtb_trim_many, one GPU lane per ion) and its collision events are kept until the
  // caller asks for that ion.  An app's queue loop costs one launch per generation of a cascade instead of one per ion.
  struct Followed
  {
    mtb_ion start;                // the ion as it was when it was followed (a changed ion is followed again)
    std::vector<mtb_event> events;
  };
  std::unordered_map<const IonBase *, Followed> _followed;
  bool followQueued(IonBase * pka, const mtb_ion & ion, std::queue<IonBase *> & recoils);
  std::vector<mtb_event> _batch_events;
  std::vector<uint32_t> _batch_counts;
  unsigned long long _seen_vac, _seen_steps;
  double _seen_eel, _seen_enuc;
};

class TrimPrimaries : public TrimBase
{
public:
  TrimPrimaries(SimconfType * simconf, SampleBase * sample) : TrimBase(simconf, sample) {}

protected:
  virtual int maxGen() const { return 1; }
  virtual bool followRecoil() { return _recoil->_gen < maxGen(); }
  virtual void vacancyCreation();
  virtual void deviceHooks(DeviceHooks & h) const;
};

class TrimRecoils : public TrimPrimaries
{
public:
  TrimRecoils(SimconfType * simconf, SampleBase * sample) : TrimPrimaries(simconf, sample) {}

protected:
  virtual int maxGen() const { return 2; }
};

class TrimHistory : public TrimBase
{
public:
  TrimHistory(SimconfType * simconf, SampleBase * sample) : TrimBase(simconf, sample) {}
  const std::vector<Point> & getHistory() { return _pos_hist; }

protected:
  virtual bool followRecoil()
  {
    _pos_hist.push_back(_pka->_pos);
    return true;
  }
  std::vector<Point> _pos_hist;
};

class TrimDefectLog : public TrimBase
{
public:
  TrimDefectLog(SimconfType * simconf, SampleBase * sample, std::ostream & os) : TrimBase(simconf, sample), _os(os) {}

protected:
  std::ostream & _os;
  virtual void vacancyCreation();
  virtual void checkPKAState();
};

class TrimVacMap : public TrimBase
{
  static const int mx = MTB_VMAP_NX, my = MTB_VMAP_NY;

public:
  TrimVacMap(SimconfType * simconf, SampleBase * sample, int z1, int z2, int z3 = -1);
  int vmap[mx][my][3];

protected:
  int _z1, _z2, _z3;
  virtual void vacancyCreation();
  virtual void deviceHooks(DeviceHooks & h) const;
  virtual void collectDeviceTallies();
};

class TrimPhononOut : public TrimBase
{
public:
  TrimPhononOut(SimconfType * simconf, SampleBase * sample, std::ostream & os) : TrimBase(simconf, sample), _os(os) {}

protected:
  std::ostream & _os;
  virtual void checkPKAState();
  virtual void dissipateRecoilEnergy();
  virtual bool followRecoil();
  /// energy totals only (the per-event lines need trim()); known to the device
  virtual void deviceHooks(DeviceHooks & h) const;
};

// ---- invert.h -----------------------------------------------------------------------------
class Inverter
{
public:
  Inverter() : maxx(0.0), maxf(0.0), tol(1e-13) {}
  virtual ~Inverter() {}
  Real x(Real f) const;

protected:
  virtual Real f(Real x) const = 0;
  Real maxx, maxf, tol;
};

class MassInverter : public Inverter
{
public:
  MassInverter();

protected:
  virtual Real f(Real x) const;
};

class EnergyInverter : public Inverter
{
public:
  EnergyInverter();
  void setMass(Real A);

protected:
  virtual Real f(Real x) const;

private:
  Real _A;
};

} // namespace MyTRIM_NS

// ---- apps/include: threaded tallies (global namespace, as in the reference) ----------------
class ThreadedTrimBase : public MyTRIM_NS::TrimBase
{
public:
  ThreadedTrimBase(MyTRIM_NS::SimconfType * simconf, MyTRIM_NS::SampleBase * sample)
    : MyTRIM_NS::TrimBase(simconf, sample), _primaries_only(false)
  {
  }
  virtual bool followRecoil() { return !_primaries_only; }
  virtual void threadJoin(const ThreadedTrimBase &) = 0;
  bool _primaries_only;
};

class TrimVacCount : public ThreadedTrimBase
{
public:
  TrimVacCount(MyTRIM_NS::SimconfType * simconf, MyTRIM_NS::SampleBase * sample);
  const std::vector<unsigned int> & vacancies() const { return _vac_bin; }
  const std::vector<unsigned int> & replacements() const { return _repl_bin; }

protected:
  virtual void vacancyCreation();
  virtual void replacementCollision();
  virtual void threadJoin(const ThreadedTrimBase & ttb);
  virtual void writeOutput();
  virtual void deviceHooks(MyTRIM_NS::DeviceHooks & h) const;
  virtual void collectDeviceTallies();
  virtual void resetDeviceBaselines() { _dev_vac.clear(); _dev_repl.clear(); }

private:
  std::vector<unsigned int> _vac_bin, _repl_bin;
  std::vector<unsigned int> _dev_vac, _dev_repl; // device totals already merged
};

class TrimVacEnergyCount : public ThreadedTrimBase
{
public:
  TrimVacEnergyCount(MyTRIM_NS::SimconfType * simconf, MyTRIM_NS::SampleBase * sample);

protected:
  virtual void vacancyCreation();
  virtual void threadJoin(const ThreadedTrimBase & ttb);
  virtual void writeOutput();
  virtual void deviceHooks(MyTRIM_NS::DeviceHooks & h) const;
  virtual void collectDeviceTallies();
  virtual void resetDeviceBaselines() { _dev_evac.clear(); }

private:
  std::vector<std::vector<unsigned int>> _evac_bin, _dev_evac;
};

class TrimRange : public ThreadedTrimBase
{
public:
  TrimRange(MyTRIM_NS::SimconfType * simconf, MyTRIM_NS::SampleBase * sample);

protected:
  virtual void vacancyCreation();
  virtual void dissipateRecoilEnergy();
  virtual bool followRecoil() { return _recoil->_gen < 1; }
  virtual void threadJoin(const ThreadedTrimBase & ttb);
  virtual void writeOutput();
  virtual void deviceHooks(MyTRIM_NS::DeviceHooks & h) const;
  virtual void collectDeviceTallies();
  virtual void resetDeviceBaselines() { _dev_seen = 0; }

private:
  std::vector<std::vector<Real>> _range;
  size_t _dev_seen;
};

#endif
