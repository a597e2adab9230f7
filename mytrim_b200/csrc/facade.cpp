// facade.cpp — host side of include/mytrim/mytrim.h: the MyTRIM plugin surface over the C ABI.
//
// Everything here is set-up, flattening and hook replay; the transport itself happens in
// mtb_engine.cu.  Functions cite the reference code whose observable behaviour they keep.
#include "mytrim/mytrim.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <sstream>
#include <typeinfo>

#include "mtb_tables.h"

namespace MyTRIM_NS
{

SimconfType * simconf = nullptr;

// ---------------------------------------------------------------------------------------------
// SimconfType — simconf.h:47-128, simconf.C:36-223
// ---------------------------------------------------------------------------------------------
SimconfType::ScoefLine::ScoefLine()
  : mm1(0), m1(0), mnat(0), rho(0), atrho(0), vfermi(0), heat(0), lfctr(0), pcoef(8, 0.0), ehigh(4, 0.0),
    screen(19, 0.0), fermicorr(15, 0.0)
{
}

SimconfType::SimconfType(unsigned int seed_value)
  : _id(0), scoef(_rows), _rng(new std::mt19937(seed_value)), _uniform(0, 1), _uniform_int(0, 65535),
    _philox_key(seed_value), _stream(0)
{
  const char * env = std::getenv("MYTRIM_DATADIR");
  _data_dir = env ? env : "";
  ed = 25.0;
  tmin = 0.2;
  tau = 0.0;
  da = 3.0;
  cw = 0.001;
  fullTraj = false;
  vacancies_created = 0;
  EelTotal = 0.0;
  EnucTotal = 0.0;
  setLengthScale(1.0);
  std::memset(snuc, 0, sizeof(snuc));
  loadTables();
}

SimconfType::~SimconfType() {}

void
SimconfType::seed(unsigned int seed_value)
{
  _rng.reset(new std::mt19937(seed_value));
  _uniform.reset();
  _uniform_int.reset();
  _philox_key = seed_value;
  _stream = 0;
}

void
SimconfType::setLengthScale(Real l)
{
  _length_scale = l;
  _area_scale = l * l;
  _volume_scale = l * l; // sic, simconf.C:78
}

namespace
{
bool
readLine(std::ifstream & f)
{
  std::string s;
  return static_cast<bool>(std::getline(f, s));
}
} // namespace

// The built-in ZBL-85 table covers every column the library reads on the hot path (SURVEY.md §8a
// a4/a10).  If MYTRIM_DATADIR points at a reference data directory, SCOEF.95A/B, SLFCTR, ELNAME
// and SNUC03 are parsed from there instead (simconf.C:81-137) so user-modified tables keep working.
void
SimconfType::loadTables()
{
  const mtb::ZblRow * rows = mtb::builtin_zbl();
  for (unsigned int i = 0; i < _rows; ++i)
  {
    ScoefLine & s = scoef[i];
    s.mm1 = rows[i].mm1;
    s.m1 = rows[i].m1;
    s.mnat = rows[i].mnat;
    s.rho = rows[i].rho;
    s.atrho = rows[i].atrho;
    s.vfermi = rows[i].vfermi;
    s.heat = rows[i].heat;
    s.lfctr = rows[i].lfctr;
    for (int k = 0; k < 8; ++k)
      s.pcoef[k] = rows[i].pcoef[k];
  }
  if (_data_dir.empty())
    return;

  std::ifstream a((_data_dir + "/SCOEF.95A").c_str()), b((_data_dir + "/SCOEF.95B").c_str()),
      l((_data_dir + "/SLFCTR.dat").c_str()), e((_data_dir + "/ELNAME.dat").c_str()),
      n((_data_dir + "/SNUC03.dat").c_str());
  if (!a || !l)
  {
    std::cerr << "MYTRIM_DATADIR=" << _data_dir << " has no SCOEF.95A/SLFCTR.dat; using built-in ZBL tables\n";
    return;
  }
  readLine(a);
  readLine(a);
  readLine(l);
  if (b)
  {
    readLine(b);
    readLine(b);
  }
  for (unsigned int i = 0; i < _rows; ++i)
  {
    ScoefLine & s = scoef[i];
    int z;
    a >> z >> s.mm1 >> s.m1 >> s.mnat >> s.rho >> s.atrho >> s.vfermi >> s.heat;
    for (int k = 0; k < 8; ++k)
      a >> s.pcoef[k];
    l >> z >> s.lfctr;
    if (b)
    {
      for (auto & v : s.ehigh)
        b >> v;
      for (auto & v : s.screen)
        b >> v;
      for (auto & v : s.fermicorr)
        b >> v;
    }
    if (e)
      e >> z >> s.sym >> s.name;
    if (!a || !l)
    {
      throw EngineError("Error reading ZBL tables from " + _data_dir); // the reference exits here (simconf.C:139-148)
    }
  }
  if (n)
    for (int i = 0; i < 92; ++i)
      for (int j = i; j < 92; ++j)
      {
        int d1, d2;
        n >> d1 >> d2 >> snuc[j][i][0] >> snuc[j][i][1] >> snuc[j][i][2] >> snuc[j][i][3];
        for (int k = 0; k < 4; ++k)
          snuc[i][j][k] = snuc[j][i][k];
      }
}

// ---------------------------------------------------------------------------------------------
// IonBase — ion.h:30-115, ion.C:26-125
// ---------------------------------------------------------------------------------------------
IonBase::IonBase() : _Z(0), _m(0), _E(0), _seed(0), _gen(0), _id(0), _tag(-1), _Ef(3.0), _state(MOVING) {}

IonBase::IonBase(IonBase * p)
  : _Z(p->_Z), _m(p->_m), _E(p->_E), _seed(0), _gen(0), _id(0), _tag(-1), _Ef(p->_Ef), _state(MOVING)
{
}

IonBase::IonBase(int Z, Real m, Real E)
  : _Z(Z), _m(m), _E(E), _seed(0), _gen(0), _id(0), _tag(-1), _Ef(3.0), _state(MOVING)
{
}

void
IonBase::setEf()
{
  _Ef = 3.0;
}

void
IonBase::parent(IonBase * p)
{
  _gen = p->_gen + 1;
  _pos = p->_pos;
  _Ef = p->_Ef;
}

IonBase *
IonBase::spawnRecoil()
{
  IonBase * r = new IonBase;
  r->parent(this);
  return r;
}

bool
IonBase::operator<(const IonBase & o) const
{
  return _Z < o._Z || (_Z == o._Z && _m < o._m);
}

std::ostream &
operator<<(std::ostream & os, const IonBase & i)
{
  return os << i._pos(0) << ' ' << i._pos(1) << ' ' << i._pos(2) << ' ' << i._Z << ' ' << i._m << ' ' << i._E << ' '
            << i._id << ' ' << i._gen << ' ' << i._tag << ' ';
}

IonBase *
IonMDTag::spawnRecoil()
{
  IonBase * r = new IonMDTag;
  r->parent(this);
  return r;
}

std::ostream &
operator<<(std::ostream & os, const IonMDTag & i)
{
  return os << static_cast<const IonBase &>(i) << i._md << ' ';
}

void
IonClock::parent(IonBase * p)
{
  IonBase::parent(p);
  IonClock * c = dynamic_cast<IonClock *>(p);
  _time = c ? c->_time : 0.0;
}

// ---------------------------------------------------------------------------------------------
// Element / MaterialBase — element.h, material.h:34-75, material.C:31-122
// ---------------------------------------------------------------------------------------------
Element::Element() : _Z(0), _m(0), _t(0), _Edisp(25.0), _Elbind(3.0), my(0), ec(0), ai(0), fi(0) {}

MaterialBase::MaterialBase(SimconfType * sc, Real rho)
  : _rho(rho), _am(0), _az(0), _arho(0), mu(0), a(0), f(0), epsdg(0), fd(0), kd(0), pmax(0), _tag(-1), _dirty(true),
    _simconf(sc), _engine(nullptr), _engine_rho(0)
{
}

MaterialBase::~MaterialBase()
{
  if (_engine)
    mtb_destroy(_engine);
}

void
MaterialBase::prepare()
{
  Real total = 0.0;
  for (auto & e : _element)
  {
    if (e._t < 0.0)
      e._t = 0.0;
    total += e._t;
  }
  _am = 0.0;
  _az = 0.0;
  for (auto & e : _element)
  {
    e._t /= total;
    _am += e._m * e._t;
    _az += Real(e._Z) * e._t;
  }
  _arho = _rho * 0.6022 / _am;
}

// Only set-up information for callers that read these public fields; the kernels derive the same
// quantities per collision from (Z1, m1) in registers.
void
MaterialBase::average(const IonBase * pka)
{
  const Real z1 = Real(pka->_Z), m1 = pka->_m;
  const Real z1p = std::pow(z1, 0.23);
  mu = m1 / _am;
  a = .5292 * .8853 / (z1p + std::pow(_az, 0.23));
  f = a * _am / (_az * z1 * 14.4 * (m1 + _am));
  epsdg = _simconf->tmin * f * Utility::pow<2>(1.0 + mu) / (4.0 * mu);
  fd = std::pow(0.01 * _az, -7.0 / 3.0);
  kd = std::pow(0.1334 * _az, 2.0 / 3.0) / std::sqrt(_am);
  for (auto & e : _element)
  {
    e.my = m1 / e._m;
    e.ec = 4.0 * e.my / Utility::pow<2>(1.0 + e.my);
    e.ai = .5292 * .8853 / (z1p + std::pow(Real(e._Z), 0.23));
    e.fi = e.ai * e._m / (z1 * Real(e._Z) * 14.4 * (m1 + e._m));
  }
  _dirty = false;
}

namespace
{
void
fillElements(const std::vector<Element> & in, std::vector<mtb_element> & out)
{
  for (const auto & e : in)
  {
    mtb_element m;
    std::memset(&m, 0, sizeof(m));
    m.Z = e._Z;
    m.m = e._m;
    m.t = e._t;
    m.Edisp = e._Edisp;
    m.Elbind = e._Elbind;
    out.push_back(m);
  }
}

void
pushTables(mtb_handle * h, SimconfType * sc)
{
  std::vector<double> pcoef(MTB_NZ * 8), vf(MTB_NZ), lf(MTB_NZ), mm1(MTB_NZ);
  for (int z = 0; z < MTB_NZ; ++z)
  {
    for (int k = 0; k < 8; ++k)
      pcoef[8 * z + k] = sc->scoef[z].pcoef[k];
    vf[z] = sc->scoef[z].vfermi;
    lf[z] = sc->scoef[z].lfctr;
    mm1[z] = sc->scoef[z].mm1;
  }
  mtb_set_tables(h, pcoef.data(), vf.data(), lf.data(), mm1.data());
}
} // namespace

Real
MaterialBase::getrstop(const IonBase * pka)
{
  bool same = _engine && _engine_rho == _rho && _engine_elements.size() == _element.size();
  for (size_t i = 0; same && i < _element.size(); ++i)
    same = _engine_elements[i]._Z == _element[i]._Z && _engine_elements[i]._m == _element[i]._m &&
           _engine_elements[i]._t == _element[i]._t;
  if (!same)
  {
    if (_engine)
      mtb_destroy(_engine);
    _engine = nullptr;
    mtb_config cfg;
    mtb_default_config(&cfg);
    cfg.device = _simconf->device;
    if (mtb_create(&cfg, &_engine) != MTB_OK)
    {
      throw EngineError(std::string("MaterialBase::getrstop: ") + mtb_last_error());
    }
    pushTables(_engine, _simconf);
    std::vector<mtb_element> els;
    fillElements(_element, els);
    mtb_material m;
    std::memset(&m, 0, sizeof(m));
    m.rho = _rho;
    m.tag = _tag;
    m.n_elements = (int)els.size();
    m.first_element = 0;
    if (mtb_set_materials(_engine, 1, &m, (int)els.size(), els.data()) != MTB_OK)
    {
      throw EngineError(std::string("MaterialBase::getrstop: ") + mtb_last_error());
    }
    _engine_elements = _element;
    _engine_rho = _rho;
  }
  const int32_t Z = pka->_Z;
  const double m1 = pka->_m, E = pka->_E;
  double out = 0.0;
  if (mtb_stopping(_engine, 0, 1, &Z, &m1, &E, &out) != MTB_OK)
  {
    throw EngineError(std::string("MaterialBase::getrstop: ") + mtb_last_error());
  }
  return out;
}

Real
MaterialBase::getDrstopDcomp(const IonBase * pka, const Element & component)
{
  // material.C:124-131: stopping cross-section of the matching element alone
  for (auto & e : _element)
    if (component._Z == e._Z && std::abs(component._m - e._m))
    {
      MaterialBase single(_simconf, 1.0);
      Element one = e;
      one._t = 1.0;
      single._element.push_back(one);
      single.prepare();
      return single.getrstop(pka) / single._arho;
    }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// samples — sample*.h/.C
// ---------------------------------------------------------------------------------------------
SampleBase::SampleBase(Real x, Real y, Real z)
{
  w[0] = x;
  w[1] = y;
  w[2] = z;
  bc[0] = bc[1] = bc[2] = PBC;
}

void
SampleBase::averages(const IonBase * pka)
{
  for (auto * m : material)
    m->average(pka);
}

Real
SampleBase::rangeMaterial(Point &, Point &)
{
  return 100000.0;
}

bool
SampleBase::describe(mtb_geometry &, std::vector<double> &) const
{
  return false;
}

void
SampleBase::describeBox(mtb_geometry & g, int kind) const
{
  std::memset(&g, 0, sizeof(g));
  g.kind = kind;
  for (int i = 0; i < 3; ++i)
  {
    g.w[i] = w[i];
    g.bc[i] = bc[i] == PBC ? MTB_BC_PBC : (bc[i] == INF ? MTB_BC_INF : MTB_BC_CUT);
  }
}

MaterialBase *
SampleSolid::lookupMaterial(Point &)
{
  return material[0];
}

bool
SampleSolid::describe(mtb_geometry & g, std::vector<double> &) const
{
  describeBox(g, MTB_GEOM_SOLID);
  return true;
}

int
SampleLayers::lookupLayer(Point & pos)
{
  Real depth = 0.0;
  size_t i = 0;
  while (i < layerThickness.size())
  {
    depth += layerThickness[i];
    if (pos(0) < depth)
      break;
    ++i;
  }
  return (int)std::min(i, material.size() - 1);
}

MaterialBase *
SampleLayers::lookupMaterial(Point & pos)
{
  return material[lookupLayer(pos)];
}

Real
SampleLayers::rangeMaterial(Point & pos, Point & dir)
{
  // distance to the next layer interface along dir (sample_layers.C:51-92); unused by the
  // transport (RANGECORRECT is compiled out in the reference) but part of the public surface
  const Real far = 1.0e6;
  if (dir(0) == 0.0)
    return far;
  const Real eps = std::abs(1.0e-10 / dir(0));
  if (pos(0) < 0.0)
    return dir(0) < 0.0 ? far : -pos(0) / dir(0) + eps;
  Real lo = 0.0;
  for (Real t : layerThickness)
  {
    if (pos(0) >= lo && pos(0) < lo + t)
      return ((dir(0) < 0 ? lo : lo + t) - pos(0)) / dir(0) + eps;
    lo += t;
  }
  return dir(0) > 0.0 ? far : (lo - pos(0)) / dir(0) + eps;
}

bool
SampleLayers::describe(mtb_geometry & g, std::vector<double> & storage) const
{
  describeBox(g, MTB_GEOM_LAYERS);
  storage.assign(layerThickness.begin(), layerThickness.end());
  g.n_layers = (int)storage.size();
  g.layer_thickness = storage.data();
  return !storage.empty();
}

SampleWire::SampleWire(Real x, Real y, Real z) : SampleBase(x, y, z)
{
  bc[0] = CUT;
  bc[1] = CUT;
}

MaterialBase *
SampleWire::lookupMaterial(Point & pos)
{
  const Real x = (pos(0) / w[0]) * 2.0 - 1.0, y = (pos(1) / w[1]) * 2.0 - 1.0;
  return (x * x + y * y) > 1.0 ? nullptr : material[0];
}

bool
SampleWire::describe(mtb_geometry & g, std::vector<double> &) const
{
  describeBox(g, MTB_GEOM_WIRE);
  return true;
}

SampleBurriedWire::SampleBurriedWire(Real x, Real y, Real z) : SampleWire(x, y, z)
{
  bc[0] = bc[1] = bc[2] = INF;
}

MaterialBase *
SampleBurriedWire::lookupMaterial(Point & pos)
{
  if (pos(2) < 0.0 && pos(2) >= -250.0)
    return material[1];
  if (pos(2) > w[2] || pos(2) < -250.0)
    return nullptr;
  MaterialBase * wire = SampleWire::lookupMaterial(pos);
  return wire ? wire : material[1];
}

bool
SampleBurriedWire::describe(mtb_geometry & g, std::vector<double> &) const
{
  describeBox(g, MTB_GEOM_BURIED_WIRE);
  return true;
}

sampleClusters::sampleClusters(Real x, Real y, Real z) : SampleBase(x, y, z), sd(0), sh(nullptr), cl(nullptr), cn(0), cnm(0), cmr(0)
{
  for (int i = 0; i < 4; ++i)
    c[i] = nullptr;
  for (int i = 0; i < 3; ++i)
  {
    kd[i] = 0;
    kn[i] = 0;
  }
}

sampleClusters::~sampleClusters()
{
  std::free(sh);
  std::free(cl);
  for (int i = 0; i < 4; ++i)
    std::free(c[i]);
}

MaterialBase *
sampleClusters::lookupMaterial(Point & pos)
{
  const int l = lookupCluster(pos, 0.0);
  if (l == -2)
    return nullptr;
  if (l == -1)
    return material[0];
  material[1]->_tag = l;
  return material[1];
}

int
sampleClusters::lookupCluster(Point & pos, Real dr)
{
  int lo[3], hi[3];
  for (int i = 0; i < 3; ++i)
  {
    int cell = (int)std::floor((pos(i) * kn[i]) / w[i]);
    if (pos(i) < 0.0 || pos(i) >= w[i])
    {
      if (bc[i] == CUT)
        return -2;
      if (bc[i] == INF)
        return -1;
      cell %= kn[i];
      if (cell < 0)
        cell += kn[i];
    }
    const int span = int((cmr + dr) / kd[i]) + 1;
    lo[i] = cell - span;
    hi[i] = cell + span;
    if (bc[i] != PBC)
    {
      lo[i] = std::max(lo[i], 0);
      hi[i] = std::min(hi[i], kn[i] - 1);
    }
  }
  auto wrap = [](int v, int n) { v %= n; return v < 0 ? v + n : v; };
  for (int ix = lo[0]; ix <= hi[0]; ++ix)
    for (int iy = lo[1]; iy <= hi[1]; ++iy)
      for (int iz = lo[2]; iz <= hi[2]; ++iz)
      {
        int l = sh[wrap(ix, kn[0]) + kn[0] * (wrap(iy, kn[1]) + kn[1] * wrap(iz, kn[2]))];
        for (; l >= 0; l = cl[l])
        {
          Real r2 = 0.0;
          for (int i = 0; i < 3; ++i)
          {
            Real d = pos(i) - c[i][l];
            if (bc[i] == PBC)
              d -= ::round(d / w[i]) * w[i];
            r2 += d * d;
          }
          if (r2 < sqr(c[3][l] + dr))
            return l;
        }
      }
  return -1;
}

void
sampleClusters::initSpatialhash(int x, int y, int z)
{
  kn[0] = x;
  kn[1] = y;
  kn[2] = z;
  std::free(sh);
  sh = (int *)std::malloc(sizeof(int) * (size_t)x * y * z);
  clearSpatialHash();
  sd = 0.0;
  for (int i = 0; i < 3; ++i)
  {
    kd[i] = w[i] / Real(kn[i]);
    sd += kd[i];
  }
  sd = 0.5 * std::sqrt(sd);
  cmr = 0.0;
}

void
sampleClusters::clearSpatialHash()
{
  std::fill(sh, sh + (size_t)kn[0] * kn[1] * kn[2], -1);
}

void
sampleClusters::reallocClusters(int n)
{
  if (n <= cnm)
    return;
  cl = (int *)std::realloc(cl, sizeof(int) * n);
  for (int i = 0; i < 4; ++i)
    c[i] = (Real *)std::realloc(c[i], sizeof(Real) * n);
  std::fill(cl + cnm, cl + n, -1);
  cnm = n;
}

void
sampleClusters::clearClusters()
{
  cn = 0;
  clearSpatialHash();
  std::fill(cl, cl + cnm, -1);
}

void
sampleClusters::addCluster(Real x, Real y, Real z, Real r)
{
  if (cn >= cnm)
    reallocClusters(cnm + cnm / 10 + 10);
  const Real p[3] = {x, y, z};
  int cell[3];
  for (int i = 0; i < 3; ++i)
  {
    c[i][cn] = p[i];
    cell[i] = int(std::floor((p[i] * kn[i]) / w[i])) % kn[i];
    if (cell[i] < 0)
      cell[i] += kn[i];
  }
  c[3][cn] = r;
  int * slot = &sh[cell[0] + kn[0] * (cell[1] + kn[1] * cell[2])];
  while (*slot >= 0)
    slot = &cl[*slot]; // append at the tail of the cell's list, as the reference does
  *slot = cn;
  cl[cn] = -1;
  cmr = std::max(cmr, r);
  ++cn;
}

void
sampleClusters::addRandomClusters(unsigned int n, Real r, Real dr, SimconfType * sc)
{
  reallocClusters(n + n / 10);
  for (unsigned int i = 0; i < n; ++i)
    for (;;)
    {
      Point p;
      for (int j = 0; j < 3; ++j)
        p(j) = sc->drand() * w[j];
      if (lookupCluster(p, dr + r) == -1)
      {
        addCluster(p(0), p(1), p(2), r);
        break;
      }
    }
}

bool
sampleClusters::describe(mtb_geometry & g, std::vector<double> & storage) const
{
  describeBox(g, MTB_GEOM_CLUSTERS);
  if (!sh)
    return false;
  storage.resize(4 * (size_t)cn);
  for (int i = 0; i < cn; ++i)
    for (int k = 0; k < 4; ++k)
      storage[4 * i + k] = c[k][i];
  for (int i = 0; i < 3; ++i)
    g.kn[i] = kn[i];
  g.n_clusters = cn;
  g.cluster_xyzr = storage.data();
  return true;
}

// ---------------------------------------------------------------------------------------------
// TrimBase — trim.h:36-108, trim.C:35-443
// ---------------------------------------------------------------------------------------------
TrimBase::TrimBase(SimconfType * sc, SampleBase * sample)
  : _potential(UNIVERSAL), _simconf(sc), _sample(sample), _pka(nullptr), _recoil(nullptr), _material(nullptr),
    _element(nullptr), recoil_queue_ptr(nullptr), terminate(false), _ls(0), _dee(0), _den(0), _base_name("mytrim"),
    _engine(nullptr), _engine_single(nullptr), _engine_fp(0), _engine_single_fp(0), _seen_vac(0), _seen_steps(0),
    _seen_eel(0), _seen_enuc(0)
{
}

TrimBase::~TrimBase()
{
  if (_engine)
    mtb_destroy(_engine);
  if (_engine_single)
    mtb_destroy(_engine_single);
}

void
TrimBase::invalidateEngine()
{
  _engine_fp = _engine_single_fp = 0;
}

namespace
{
inline void
fnv(unsigned long long & h, const void * p, size_t n)
{
  const unsigned char * b = static_cast<const unsigned char *>(p);
  for (size_t i = 0; i < n; ++i)
    h = (h ^ b[i]) * 1099511628211ull;
}
} // namespace

unsigned long long
TrimBase::configFingerprint(bool batch, const mtb_config & cfg, const std::vector<mtb_material> & mats,
                            const std::vector<mtb_element> & els, const mtb_geometry & g,
                            const std::vector<double> & storage) const
{
  unsigned long long h = 1469598103934665603ull ^ (batch ? 1u : 0u);
  fnv(h, &cfg, sizeof(cfg));
  if (!mats.empty())
    fnv(h, mats.data(), mats.size() * sizeof(mtb_material));
  if (!els.empty())
    fnv(h, els.data(), els.size() * sizeof(mtb_element));
  mtb_geometry gg = g; // the pointers address `storage`, whose content is hashed instead
  gg.layer_thickness = nullptr;
  gg.cluster_xyzr = nullptr;
  fnv(h, &gg, sizeof(gg));
  if (!storage.empty())
    fnv(h, storage.data(), storage.size() * sizeof(double));
  return h ? h : 1ull;
}

bool
TrimBase::followRecoil()
{
  return true;
}

void
TrimBase::vacancyCreation()
{
  _simconf->vacancies_created++;
}

void
TrimBase::deviceHooks(DeviceHooks & h) const
{
  // Only plain TrimBase is known here.  A subclass that inherits this implementation (TrimHistory,
  // TrimDefectLog, any user class) has hooks the device cannot see: it must override deviceHooks()
  // or use trim() per ion.
  h.known = typeid(*this) == typeid(TrimBase);
}

mtb_handle *
TrimBase::engine()
{
  return _engine;
}

// (Re)creates the engine for this Trim object: flattens SimconfType, the sample's materials and
// geometry and (for trimBatch) the subclass' hook description.
bool
TrimBase::ensureEngine(bool batch)
{
  mtb_handle *& eng = batch ? _engine : _engine_single;
  unsigned long long & eng_fp = batch ? _engine_fp : _engine_single_fp;

  mtb_config cfg;
  std::memset(&cfg, 0, sizeof(cfg)); // padding bytes are part of the fingerprint
  mtb_default_config(&cfg);
  cfg.tmin = _simconf->tmin;
  cfg.tau = _simconf->tau;
  cfg.cw = _simconf->cw;
  cfg.length_scale = _simconf->lengthScale();
  cfg.potential = (int)_potential;
  cfg.device = _simconf->device;
  if (batch)
  {
    DeviceHooks h;
    deviceHooks(h);
    if (!h.known)
    {
      _error = "this TrimBase subclass has host-only hooks (deviceHooks() reports unknown): use trim() per ion";
      return false;
    }
    cfg.follow = h.follow;
    cfg.follow_max_gen = h.follow_max_gen;
    cfg.vacancy_model = h.vacancy_model;
    cfg.tally_mask = h.tally_mask | MTB_TALLY_RECORDS;
    for (int i = 0; i < 3; ++i)
      cfg.vmap_z[i] = h.vmap_z[i];
    cfg.ionlog_z = h.ionlog_z;
    cfg.hist_bins = h.hist_bins;
    cfg.ionlog_capacity = h.ionlog_capacity;
    cfg.range_capacity = h.range_capacity;
  }

  std::vector<mtb_material> mats;
  std::vector<mtb_element> els;
  for (auto * m : _sample->material)
  {
    mtb_material mm;
    std::memset(&mm, 0, sizeof(mm));
    mm.rho = m->_rho;
    mm.tag = m->_tag;
    mm.first_element = (int)els.size();
    mm.n_elements = (int)m->_element.size();
    fillElements(m->_element, els);
    mats.push_back(mm);
  }
  std::vector<double> storage;
  mtb_geometry g;
  std::memset(&g, 0, sizeof(g));
  if (!_sample->describe(g, storage))
  {
    _error = "this SampleBase subclass has a host-only lookupMaterial(): the device needs describe()";
    return false;
  }

  // the reference reads SimconfType, _potential, the materials and the sample on every trim() call: rebuild the
  // engine when any of them changed since the snapshot
  const unsigned long long fp = configFingerprint(batch, cfg, mats, els, g, storage);
  if (eng && eng_fp == fp)
    return true;
  if (eng)
    mtb_destroy(eng);
  eng = nullptr;
  eng_fp = 0;
  if (batch)
  {
    // a new batch engine starts its tallies at zero
    _seen_vac = _seen_steps = 0;
    _seen_eel = _seen_enuc = 0.0;
    resetDeviceBaselines();
  }

  if (mtb_create(&cfg, &eng) != MTB_OK)
  {
    _error = mtb_last_error();
    eng = nullptr;
    return false;
  }
  pushTables(eng, _simconf);
  if (mtb_set_materials(eng, (int)mats.size(), mats.data(), (int)els.size(), els.data()) != MTB_OK ||
      mtb_set_geometry(eng, &g) != MTB_OK)
  {
    _error = mtb_last_error();
    mtb_destroy(eng);
    eng = nullptr;
    return false;
  }
  eng_fp = fp;
  return true;
}

namespace
{
void
ionToAbi(const IonBase * in, mtb_ion & o)
{
  std::memset(&o, 0, sizeof(o));
  for (int i = 0; i < 3; ++i)
  {
    o.pos[i] = in->_pos(i);
    o.dir[i] = in->_dir(i);
  }
  o.E = in->_E;
  o.m = in->_m;
  o.Ef = in->_Ef;
  o.Z = in->_Z;
  o.gen = in->_gen;
  o.tag = in->_tag;
  o.seed = in->_seed;
}
} // namespace

namespace
{
// the container behind a std::queue (its protected member `c`), read-only
template <class T>
const std::deque<T> &
queueContainer(const std::queue<T> & q)
{
  struct Access : std::queue<T>
  {
    static const std::deque<T> & get(const std::queue<T> & q) { return q.*(&Access::c); }
  };
  return Access::get(q);
}
} // namespace

// Follows `pka` and every ion that waits in the caller's queue and has not been followed yet in one launch
// (mtb_trim_many), ions with long trajectories in further launches with buffers of the size the first one reported.
bool
TrimBase::followQueued(IonBase * pka, const mtb_ion & ion, std::queue<IonBase *> & recoils)
{
  const size_t kMaxBatch = 8192, kFirstEvents = 32, kMaxEventsPerLaunch = 1u << 20;
  // ions that were followed ahead of time but never asked for (an app that drops part of its queue) must not pile up:
  // forgetting them only costs a repeated launch if they are asked for after all
  if (_followed.size() > 4 * kMaxBatch)
    _followed.clear();
  std::vector<const IonBase *> who(1, pka);
  std::vector<mtb_ion> ions(1, ion);
  for (IonBase * q : queueContainer(recoils))
  {
    if (who.size() >= kMaxBatch)
      break;
    if (q == pka || !q)
      continue;
    mtb_ion a;
    ionToAbi(q, a);
    auto it = _followed.find(q);
    if (it != _followed.end())
    {
      if (std::memcmp(&it->second.start, &a, sizeof(a)) == 0)
        continue;
      _followed.erase(it);
    }
    who.push_back(q);
    ions.push_back(a);
  }
  const size_t n = who.size();
  const uint64_t first_uid = _simconf->nextStreamId(n);
  const std::vector<mtb_ion> start = ions;
  _batch_events.resize(n * kFirstEvents);
  _batch_counts.resize(n);
  if (mtb_trim_many(_engine_single, n, ions.data(), _simconf->philoxKey(), first_uid, nullptr, nullptr, _batch_events.data(),
                    kFirstEvents, _batch_counts.data()) != MTB_OK)
  {
    _error = mtb_last_error();
    return false;
  }
  std::vector<size_t> longer; // ions with more collisions than the first buffer holds
  for (size_t i = 0; i < n; ++i)
  {
    if (_batch_counts[i] > kFirstEvents)
    {
      longer.push_back(i);
      continue;
    }
    Followed & f = _followed[who[i]];
    f.start = start[i];
    f.events.assign(_batch_events.begin() + i * kFirstEvents, _batch_events.begin() + i * kFirstEvents + _batch_counts[i]);
  }
  // the long ones again, grouped so that a launch holds at most kMaxEventsPerLaunch events; same stream ids, so the
  // trajectories are the ones the first launch counted
  std::sort(longer.begin(), longer.end(), [this](size_t a, size_t b) { return _batch_counts[a] < _batch_counts[b]; });
  const std::vector<uint32_t> counts = _batch_counts;
  for (size_t lo = 0; lo < longer.size();)
  {
    size_t hi = lo + 1;
    while (hi < longer.size() && (hi - lo + 1) * (size_t)counts[longer[hi]] <= kMaxEventsPerLaunch)
      ++hi;
    const size_t m = hi - lo, cap = counts[longer[hi - 1]];
    std::vector<mtb_ion> sub(m);
    std::vector<uint64_t> uids(m);
    for (size_t k = 0; k < m; ++k)
    {
      sub[k] = start[longer[lo + k]];
      uids[k] = first_uid + longer[lo + k];
    }
    _batch_events.resize(m * cap);
    _batch_counts.resize(m);
    if (mtb_trim_many(_engine_single, m, sub.data(), _simconf->philoxKey(), 0, uids.data(), nullptr, _batch_events.data(), cap,
                      _batch_counts.data()) != MTB_OK)
    {
      _error = mtb_last_error();
      return false;
    }
    for (size_t k = 0; k < m; ++k)
    {
      if (_batch_counts[k] != counts[longer[lo + k]])
      {
        _error = "event replay of an ion differs from its first pass";
        return false;
      }
      Followed & f = _followed[who[longer[lo + k]]];
      f.start = start[longer[lo + k]];
      f.events.assign(_batch_events.begin() + k * cap, _batch_events.begin() + k * cap + _batch_counts[k]);
    }
    lo = hi;
  }
  return true;
}

// One ion: the device follows it (with everything else that waits in the caller's queue, followQueued), the hooks run
// here in the reference's order
// (trim.C:357-418): followRecoil -> vacancyCreation | replacementCollision, or
// dissipateRecoilEnergy; then checkPKAState.  Recoils the hooks accept go to the caller's queue.
void
TrimBase::trim(IonBase * pka, std::queue<IonBase *> & recoils)
{
  _pka = pka;
  _pka->_state = IonBase::MOVING;
  recoil_queue_ptr = &recoils;
  if (!ensureEngine(false))
  {
    throw EngineError("TrimBase::trim: " + _error);
  }
  mtb_ion ion;
  ionToAbi(pka, ion);
  auto hit = _followed.find(pka);
  if (hit != _followed.end() && std::memcmp(&hit->second.start, &ion, sizeof(ion)) != 0)
  {
    _followed.erase(hit); // the caller changed the ion after it had been followed (or the address was re-used)
    hit = _followed.end();
  }
  if (hit == _followed.end())
  {
    if (!followQueued(pka, ion, recoils))
    {
      throw EngineError("TrimBase::trim: " + _error);
    }
    hit = _followed.find(pka);
  }
  _events.swap(hit->second.events);
  _followed.erase(hit);
  const size_t n = _events.size();

  for (size_t k = 0; k < n; ++k)
  {
    const mtb_event & ev = _events[k];
    _material = _sample->material[ev.material];
    _material->_tag = ev.material_tag;
    if (_material->_dirty)
      _material->average(_pka);
    _element = &_material->getElement(ev.element);
    _ls = ev.ls;
    _dee = ev.dee;
    _den = ev.den;
    _simconf->EelTotal += ev.dee;

    // the recoil is spawned before the projectile moves (trim.C:306-318)
    _pka->_pos = Point(ev.recoil_pos[0], ev.recoil_pos[1], ev.recoil_pos[2]);
    _recoil = _pka->spawnRecoil();
    _recoil->_dir = Point(ev.recoil_dir[0], ev.recoil_dir[1], ev.recoil_dir[2]);
    _recoil->_E = ev.recoil_E;
    _recoil->_m = _element->_m;
    _recoil->_Z = _element->_Z;
    _recoil->_state = IonBase::MOVING;

    _pka->_pos = Point(ev.pka_pos[0], ev.pka_pos[1], ev.pka_pos[2]);
    _pka->_dir = Point(ev.pka_dir[0], ev.pka_dir[1], ev.pka_dir[2]);
    _pka->_E = ev.pka_E;
    _pka->_state = ev.pka_state == MTB_LOST ? IonBase::LOST : IonBase::MOVING;

    if (_pka->_state != IonBase::LOST)
    {
      if (ev.recoil_above_threshold)
      {
        if (followRecoil())
        {
          _recoil->_tag = _material->_tag;
          _recoil->_id = _simconf->_id++;
          recoils.push(_recoil);
          if (_simconf->fullTraj)
            std::cout << "spawn " << _recoil->_id << ' ' << _pka->_id << '\n';
        }
        else
          _recoil->_state = IonBase::DELETE;
        if (ev.pka_state == MTB_MOVING)
          vacancyCreation();
        else
        {
          replacementCollision();
          _pka->_state = (IonBase::StateType)ev.pka_state;
        }
      }
      else
      {
        dissipateRecoilEnergy();
        _recoil->_state = IonBase::DELETE;
        _pka->_state = (IonBase::StateType)ev.pka_state;
      }
    }
    if (_recoil->_state == IonBase::DELETE)
      delete _recoil;
    checkPKAState();
    if (_simconf->fullTraj)
      std::cout << _pka->_state << ' ' << *_pka << '\n';
  }
  if (n == 0)
    _pka->_state = IonBase::MOVING; // started in vacuum (trim.C:80-82)
}

bool
TrimBase::trimBatch(std::vector<IonBase *> & primaries)
{
  return trimBatch(primaries, nullptr);
}

bool
TrimBase::trimBatch(std::vector<IonBase *> & primaries, std::vector<mtb_record> * records_out)
{
  if (!ensureEngine(true))
    return false;
  const size_t n = primaries.size();
  std::vector<mtb_ion> ions(n);
  for (size_t i = 0; i < n; ++i)
    ionToAbi(primaries[i], ions[i]);
  std::vector<mtb_record> local;
  std::vector<mtb_record> & rec = records_out ? *records_out : local;
  rec.resize(n);
  DeviceHooks hooks;
  deviceHooks(hooks);
  const size_t chunk = hooks.batch_chunk ? (size_t)hooks.batch_chunk : std::max<size_t>(n, 1);
  const uint64_t first = _simconf->nextStreamId(n);
  for (size_t lo = 0; lo < n || lo == 0; lo += chunk)
  {
    const size_t m = std::min(chunk, n - lo);
    if (mtb_run(_engine, m, ions.data() + lo, _simconf->philoxKey(), first + lo, rec.data() + lo) != MTB_OK)
    {
      _error = mtb_last_error();
      return false;
    }
    mtb_counters c;
    if (mtb_get_counters(_engine, &c) != MTB_OK)
    {
      _error = mtb_last_error();
      return false;
    }
    _simconf->vacancies_created += (int)(c.vacancies_created - _seen_vac);
    _simconf->EelTotal += c.EelTotal - _seen_eel;
    _simconf->EnucTotal += c.EnucTotal - _seen_enuc;
    _seen_vac = c.vacancies_created;
    _seen_eel = c.EelTotal;
    _seen_enuc = c.EnucTotal;
    _seen_steps = c.steps;
    _error.clear();
    collectDeviceTallies();
    if (!_error.empty())
      return false;
    if (n == 0)
      break;
  }
  for (size_t i = 0; i < n; ++i)
  {
    primaries[i]->_pos = Point(rec[i].pos[0], rec[i].pos[1], rec[i].pos[2]);
    primaries[i]->_E = rec[i].E;
    primaries[i]->_state = (IonBase::StateType)rec[i].state;
  }
  return true;
}

// ---- built-in subclasses (trim.h:113-226, trim.C:445-527) ----------------------------------
void
TrimPrimaries::vacancyCreation()
{
  _simconf->vacancies_created++;
  if (_recoil->_gen == maxGen())
  {
    // modified Kinchin-Pease estimate for the cascade that is not followed
    const Real ed = 0.0115 * std::pow(_material->_az, -7.0 / 3.0) * _recoil->_E;
    const Real g = 3.4008 * std::pow(ed, 1.0 / 6.0) + 0.40244 * std::pow(ed, 3.0 / 4.0) + ed;
    const Real kd = 0.1337 * std::pow(_material->_az, 2.0 / 3.0) / std::sqrt(_material->_am);
    const Real Ev = _recoil->_E / (1.0 + kd * g);
    _simconf->vacancies_created += int(0.8 * Ev / (2.0 * _element->_Edisp));
  }
}

void
TrimPrimaries::deviceHooks(DeviceHooks & h) const
{
  h.known = true;
  h.follow = MTB_FOLLOW_GEN_LT;
  h.follow_max_gen = maxGen();
  h.vacancy_model = MTB_VAC_KP;
}

void
TrimDefectLog::vacancyCreation()
{
  _os << "V " << *_recoil << '\n';
}

void
TrimDefectLog::checkPKAState()
{
  const char * tag = _pka->_state == IonBase::INTERSTITIAL     ? "I "
                     : _pka->_state == IonBase::SUBSTITUTIONAL ? "S "
                     : _pka->_state == IonBase::REPLACEMENT    ? "R "
                                                               : nullptr;
  if (tag)
    _os << tag << *_pka << '\n';
}

TrimVacMap::TrimVacMap(SimconfType * sc, SampleBase * sample, int z1, int z2, int z3)
  : TrimBase(sc, sample), _z1(z1), _z2(z2), _z3(z3)
{
  std::memset(vmap, 0, sizeof(vmap));
}

void
TrimVacMap::vacancyCreation()
{
  int x = (_recoil->_pos(0) * mx) / _sample->w[0];
  int y = (_recoil->_pos(1) * my) / _sample->w[1];
  x -= int(x / mx) * mx;
  y -= int(y / my) * my;
  if (x < 0 || y < 0)
    return; // the reference indexes out of bounds here
  const int s = _recoil->_Z == _z1 ? 0 : (_recoil->_Z == _z2 ? 1 : (_recoil->_Z == _z3 ? 2 : -1));
  if (s >= 0)
    vmap[x][y][s]++;
}

void
TrimVacMap::deviceHooks(DeviceHooks & h) const
{
  h.known = true;
  h.vacancy_model = MTB_VAC_NONE;
  h.tally_mask = MTB_TALLY_VACMAP;
  h.vmap_z[0] = _z1;
  h.vmap_z[1] = _z2;
  h.vmap_z[2] = _z3;
}

void
TrimVacMap::collectDeviceTallies()
{
  std::vector<uint64_t> v(mx * my * 3);
  if (mtb_get_vacmap(engine(), v.data()) != MTB_OK)
    return;
  for (int x = 0; x < mx; ++x)
    for (int y = 0; y < my; ++y)
      for (int s = 0; s < 3; ++s)
        vmap[x][y][s] = (int)v[(x * my + y) * 3 + s];
}

void
TrimPhononOut::checkPKAState()
{
  if (_pka->_state == IonBase::MOVING || _pka->_state == IonBase::LOST)
    return;
  _os << _pka->_E << ' ' << *_pka << '\n';
  _simconf->EnucTotal += _pka->_E;
}

void
TrimPhononOut::dissipateRecoilEnergy()
{
  const Real Edep = _recoil->_E + _element->_Elbind;
  _os << Edep << ' ' << *_recoil << '\n';
  _simconf->EnucTotal += Edep;
}

bool
TrimPhononOut::followRecoil()
{
  _os << _element->_Elbind << ' ' << *_recoil << '\n';
  _simconf->EnucTotal += _element->_Elbind;
  return true;
}

void
TrimPhononOut::deviceHooks(DeviceHooks & h) const
{
  h.known = true;
  h.tally_mask = MTB_TALLY_PHONON;
}

// ---- invert.h / invert.C -------------------------------------------------------------------
Real
Inverter::x(Real target) const
{
  Real pos = maxx / 2.0, step = maxx / 4.0;
  for (int i = 0; i < 32; ++i)
  {
    const Real val = f(pos) / maxf;
    if (std::abs(val - target) <= tol)
      break;
    pos += val > target ? -step : step;
    step *= 0.5;
  }
  return pos;
}

MassInverter::MassInverter()
{
  maxx = 235.0;
  tol = 1e-7;
  maxf = f(maxx);
}

Real
MassInverter::f(Real x) const
{
  // cumulative fission mass yield, H.R. Faust, Eur. Phys. J. A 14 (2002) 459 (single-precision erf
  // as in the reference, invert.C:47-56)
  static const Real amp[5] = {0.112798, 37.4781, 37.4781, 12.5094, 12.5094};
  static const Real off[5] = {-5.56257, -19.3772, -13.0462, -30.8853, -23.2853};
  static const Real slope[5] = {0.0471405, 0.137386, 0.137386, 0.229537, 0.229537};
  Real sum = 100.088;
  for (int i = 0; i < 5; ++i)
    sum += amp[i] * erff(off[i] + slope[i] * x);
  return sum / 200.1756;
}

EnergyInverter::EnergyInverter()
{
  maxx = 186.98;
  tol = 1e-7;
  setMass(100.0);
}

void
EnergyInverter::setMass(Real A)
{
  _A = A;
  maxf = f(maxx);
}

Real
EnergyInverter::f(Real x) const
{
  const Real x1 = x / (1.0 - _A / 234.0);
  return (-0.00014122 + (0.00014122 - 7.12299E-7 * x1) * std::exp(0.0886603 * x1)) / 127.216;
}

} // namespace MyTRIM_NS

// ---- C ABI: the fission-fragment source of the UO2 experiment (include/mytrim_b200.h) -------
extern "C" int
mtb_fission_pairs(uint32_t seed, uint64_t first_event, uint64_t n_events, const double w[3], mtb_ion * out, double * e_total)
{
  if (!w || (n_events && !out))
    return MTB_EINVAL;
  std::mt19937 rng(seed);
  std::uniform_real_distribution<double> uniform(0, 1);
  MyTRIM_NS::MassInverter mass;
  MyTRIM_NS::EnergyInverter energy;
  double sum = 0.0;
  uint64_t degenerate = 0;
  for (uint64_t ev = 0; ev < first_event + n_events; ++ev)
  {
    // same draw order as apps/mytrim_uo2.C:226-262: mass, energy, direction (rejection), origin
    const double A1 = mass.x(uniform(rng)), A2 = 235.0 - A1;
    energy.setMass(A1);
    const double Etot = energy.x(uniform(rng));
    double d[3], norm;
    do
    {
      for (int i = 0; i < 3; ++i)
        d[i] = 2.0 * uniform(rng) - 1.0;
      norm = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    } while (norm <= 0.0001 || norm > 1.0);
    double pos[3];
    for (int i = 0; i < 3; ++i)
      pos[i] = uniform(rng) * w[i];
    if (ev < first_event)
      continue; // the source stream is sequential: a shard skips the events before its range
    const double E1 = Etot * A2 / (A1 + A2), E2 = Etot - E1;
    const int Z1 = (int)std::round((A1 * 92.0) / 235.0);
    const double inv = 1.0 / std::sqrt(norm);
    mtb_ion * o = out + 2 * (ev - first_event);
    std::memset(o, 0, 2 * sizeof(mtb_ion));
    for (int k = 0; k < 2; ++k)
    {
      for (int i = 0; i < 3; ++i)
      {
        o[k].pos[i] = pos[i];
        o[k].dir[i] = (k ? -1.0 : 1.0) * d[i] * inv;
      }
      o[k].E = (k ? E2 : E1) * 1.0e6;
      o[k].m = k ? A2 : A1;
      o[k].Z = k ? 92 - Z1 : Z1;
      o[k].Ef = 3.0;
      o[k].gen = 0;
      o[k].tag = -1;
      if (o[k].Z < 1 || o[k].Z > 92)
      {
        // About one draw in 1e6 falls so far into the tail of the mass distribution that the 32-step bisection of
        // Inverter::x ends at A = 235 * 2^-33 and the fragment gets Z = 0.  The reference then reads scoef[-1]
        // (material.C:163-187): undefined behaviour.  Here the fragment keeps its place in the list (its index is its
        // Philox stream) but carries no energy, so it stops where it starts; *n_degenerate counts them.
        o[k].Z = o[k].Z < 1 ? 1 : 92;
        o[k].m = std::max(o[k].m, 1.0);
        o[k].E = 0.0;
        ++degenerate;
      }
      sum += o[k].E;
    }
  }
  if (e_total)
    *e_total = sum;
  if (degenerate)
    std::fprintf(stderr, "mtb_fission_pairs: %llu fragment(s) with Z outside 1..92 emitted without energy\n",
                 (unsigned long long)degenerate);
  return MTB_OK;
}

// ---------------------------------------------------------------------------------------------
// apps/include + apps/src: the threaded tallies runmytrim uses
// ---------------------------------------------------------------------------------------------
using namespace MyTRIM_NS;

namespace
{
template <class T>
void
bump(std::vector<T> & v, int x)
{
  if (x < 0)
    return;
  if (x >= (int)v.size())
    v.resize(x + 1, 0);
  v[x]++;
}

template <class T>
void
addInto(std::vector<T> & dst, const std::vector<T> & src)
{
  if (dst.size() < src.size())
    dst.resize(src.size(), 0);
  for (size_t i = 0; i < src.size(); ++i)
    dst[i] += src[i];
}
} // namespace

TrimVacCount::TrimVacCount(SimconfType * sc, SampleBase * sample) : ThreadedTrimBase(sc, sample) {}

void
TrimVacCount::vacancyCreation()
{
  _simconf->vacancies_created++;
  bump(_vac_bin, int(_recoil->_pos(0)));
}

void
TrimVacCount::replacementCollision()
{
  bump(_repl_bin, int(_recoil->_pos(0)));
}

void
TrimVacCount::threadJoin(const ThreadedTrimBase & other)
{
  const TrimVacCount & o = static_cast<const TrimVacCount &>(other);
  addInto(_vac_bin, o._vac_bin);
  addInto(_repl_bin, o._repl_bin);
}

void
TrimVacCount::writeOutput()
{
  const size_t n = std::max(_vac_bin.size(), _repl_bin.size());
  _vac_bin.resize(n);
  _repl_bin.resize(n);
  std::ofstream out((_base_name + "_vac.dat").c_str());
  for (size_t x = 0; x < n; ++x)
    out << x << ' ' << _vac_bin[x] << ' ' << _repl_bin[x] << '\n';
}

void
TrimVacCount::deviceHooks(DeviceHooks & h) const
{
  h.known = true;
  h.follow = _primaries_only ? MTB_FOLLOW_NONE : MTB_FOLLOW_ALL;
  h.tally_mask = MTB_TALLY_VAC_DEPTH;
}

// device histograms are cumulative: merge only what is new since the last call
void
TrimVacCount::collectDeviceTallies()
{
  size_t bins = 0;
  if (mtb_hist_bins(engine(), &bins, nullptr) != MTB_OK)
    return;
  std::vector<uint64_t> v(bins), r(bins);
  size_t n = 0;
  if (mtb_get_vac_depth(engine(), v.data(), r.data(), bins, &n) != MTB_OK)
    return;
  _dev_vac.resize(std::max(_dev_vac.size(), n), 0);
  _dev_repl.resize(std::max(_dev_repl.size(), n), 0);
  if (_vac_bin.size() < n)
    _vac_bin.resize(n, 0);
  if (_repl_bin.size() < n)
    _repl_bin.resize(n, 0);
  for (size_t i = 0; i < n; ++i)
  {
    _vac_bin[i] += (unsigned int)(v[i] - _dev_vac[i]);
    _repl_bin[i] += (unsigned int)(r[i] - _dev_repl[i]);
    _dev_vac[i] = (unsigned int)v[i];
    _dev_repl[i] = (unsigned int)r[i];
  }
  // trailing all-zero bins do not exist in the reference's dynamically grown vectors
  while (!_vac_bin.empty() && _vac_bin.back() == 0)
    _vac_bin.pop_back();
  while (!_repl_bin.empty() && _repl_bin.back() == 0)
    _repl_bin.pop_back();
}

TrimVacEnergyCount::TrimVacEnergyCount(SimconfType * sc, SampleBase * sample) : ThreadedTrimBase(sc, sample) {}

void
TrimVacEnergyCount::vacancyCreation()
{
  _simconf->vacancies_created++;
  const int x = int(_recoil->_pos(0));
  if (x < 0)
    return;
  const int row = std::max(0, int(std::log(_recoil->_E)));
  if (row >= (int)_evac_bin.size())
    _evac_bin.resize(row + 1);
  bump(_evac_bin[row], x);
}

void
TrimVacEnergyCount::threadJoin(const ThreadedTrimBase & other)
{
  const TrimVacEnergyCount & o = static_cast<const TrimVacEnergyCount &>(other);
  if (_evac_bin.size() < o._evac_bin.size())
    _evac_bin.resize(o._evac_bin.size());
  for (size_t e = 0; e < o._evac_bin.size(); ++e)
    addInto(_evac_bin[e], o._evac_bin[e]);
}

void
TrimVacEnergyCount::writeOutput()
{
  std::ofstream out((_base_name + "_evac.dat").c_str());
  for (size_t e = 0; e < _evac_bin.size(); ++e)
  {
    for (size_t x = 0; x < _evac_bin[e].size(); ++x)
      out << e << ' ' << x << ' ' << _evac_bin[e][x] << '\n';
    out << '\n';
  }
}

void
TrimVacEnergyCount::deviceHooks(DeviceHooks & h) const
{
  h.known = true;
  h.follow = _primaries_only ? MTB_FOLLOW_NONE : MTB_FOLLOW_ALL;
  h.tally_mask = MTB_TALLY_VAC_ENERGY;
}

void
TrimVacEnergyCount::collectDeviceTallies()
{
  size_t bins = 0, rows = 0;
  if (mtb_hist_bins(engine(), &bins, &rows) != MTB_OK)
    return;
  std::vector<uint64_t> buf(rows * bins);
  if (mtb_get_vac_energy(engine(), buf.data(), rows, bins) != MTB_OK)
    return;
  if (_dev_evac.size() < rows)
    _dev_evac.resize(rows);
  for (size_t r = 0; r < rows; ++r)
  {
    size_t last = 0;
    for (size_t x = 0; x < bins; ++x)
      if (buf[r * bins + x])
        last = x + 1;
    if (!last)
      continue;
    if (_evac_bin.size() <= r)
      _evac_bin.resize(r + 1);
    if (_evac_bin[r].size() < last)
      _evac_bin[r].resize(last, 0);
    if (_dev_evac[r].size() < last)
      _dev_evac[r].resize(last, 0);
    for (size_t x = 0; x < last; ++x)
    {
      _evac_bin[r][x] += (unsigned int)(buf[r * bins + x] - _dev_evac[r][x]);
      _dev_evac[r][x] = (unsigned int)buf[r * bins + x];
    }
  }
}

TrimRange::TrimRange(SimconfType * sc, SampleBase * sample) : ThreadedTrimBase(sc, sample), _range(MTB_MAX_RANGE_Z), _dev_seen(0) {}

void
TrimRange::vacancyCreation()
{
  // NRT damage estimate of the sub-cascade that is not followed (apps/src/TrimRange.C:31-47)
  const Real Ed = _element->_Edisp;
  const Real ed = 0.0115 * std::pow(_recoil->_Z, -7.0 / 3.0) * _recoil->_E;
  const Real kd = 0.1337 * std::pow(_recoil->_Z, 2.0 / 3.0) / std::sqrt(_recoil->_m);
  const Real g = 3.4008 * std::pow(ed, 1.0 / 6.0) + 0.40244 * std::pow(ed, 3.0 / 4.0) + ed;
  const Real Ev = _recoil->_E / (1.0 + kd * g);
  if (Ev < Ed)
    return;
  if (Ev >= Ed / 0.4)
    _simconf->vacancies_created += Ev * 0.4 / Ed;
  else
    _simconf->vacancies_created++;
}

void
TrimRange::dissipateRecoilEnergy()
{
  _range[_recoil->_Z].push_back(_recoil->_pos(0));
}

void
TrimRange::threadJoin(const ThreadedTrimBase & other)
{
  const TrimRange & o = static_cast<const TrimRange &>(other);
  for (size_t Z = 0; Z < _range.size(); ++Z)
    _range[Z].insert(_range[Z].end(), o._range[Z].begin(), o._range[Z].end());
}

void
TrimRange::writeOutput()
{
  // one histogram column per Z with a common binning (apps/src/TrimRange.C:66-121)
  size_t most = 0;
  Real lo = 0.0, hi = 0.0;
  for (auto & list : _range)
  {
    most = std::max(most, list.size());
    for (Real x : list)
    {
      lo = std::min(lo, x);
      hi = std::max(hi, x);
    }
  }
  const Real width = hi - lo;
  const Real bin = std::min(width * 100.0 / most, width / 10.0);
  const unsigned int nbin = width / bin + 1;
  std::vector<std::pair<int, std::vector<unsigned int>>> columns;
  for (size_t Z = 0; Z < _range.size(); ++Z)
  {
    if (_range[Z].empty())
      continue;
    std::vector<unsigned int> h(nbin);
    for (Real x : _range[Z])
      h[std::floor((x - lo) / bin)]++;
    columns.push_back(std::make_pair((int)Z, h));
  }
  std::ofstream out((_base_name + "_ranges.dat").c_str());
  out << "#x";
  for (auto & c : columns)
    out << " Z" << c.first;
  out << '\n';
  for (unsigned int i = 0; i < nbin; ++i)
  {
    out << (i * bin + lo);
    for (auto & c : columns)
      out << ' ' << c.second[i];
    out << '\n';
  }
}

void
TrimRange::deviceHooks(DeviceHooks & h) const
{
  h.known = true;
  h.follow = MTB_FOLLOW_NONE; // _recoil->_gen < 1 never holds (TrimRange.h:16)
  h.vacancy_model = MTB_VAC_NRT;
  h.tally_mask = MTB_TALLY_RANGE;
  // every sub-threshold recoil of every primary is one list entry (~220 per 150 keV Cu primary): the primaries go
  // to the device in chunks and the list is drained after each
  h.batch_chunk = 8192;
  h.range_capacity = 1ull << 24;
}

void
TrimRange::collectDeviceTallies()
{
  size_t n = 0;
  mtb_get_range_list(engine(), nullptr, nullptr, 0, &n);
  if (n == 0)
    return;
  std::vector<float> x(n);
  std::vector<int32_t> Z(n);
  const int rc = mtb_get_range_list(engine(), x.data(), Z.data(), n, &n);
  if (rc != MTB_OK)
  {
    // an overflowed list would silently truncate <base>_ranges.dat
    _error = std::string("TrimRange: ") + mtb_last_error();
    return;
  }
  for (size_t i = 0; i < std::min(n, x.size()); ++i)
    if (Z[i] >= 0 && Z[i] < (int)_range.size())
      _range[Z[i]].push_back(x[i]);
  mtb_clear_lists(engine()); // drained: the next chunk starts an empty list
  _dev_seen = 0;
}
