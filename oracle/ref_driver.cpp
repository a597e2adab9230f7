// ref_driver.cpp — drives the UNMODIFIED reference library (compiled from /root/reference by
// oracle/Makefile into oracle/_ref/) so that its outputs can pin the oracle and serve as the
// CPU baseline.  TEST INFRASTRUCTURE ONLY.
//
// jsoncpp is not installed in this image, so apps/runmytrim.C cannot be built as is; this driver
// constructs the same objects runmytrim does (apps/runmytrim.C:60-93, 183-331: one private
// {SimconfType, SampleLayers, materials, Trim*} per thread, primaries dealt round-robin, every
// primary reseeds its thread's RNG) from a small line-oriented description on stdin and adds
// per-primary records.  All physics runs inside the reference's own TrimBase::trim().
//
// stdin commands (one per line):
//   ion Z m E [Ef]            primary species
//   n N                       number of primaries
//   threads T
//   seeds <file>              binary uint32[N] per-primary seeds (default: irand() of `master`)
//   master S                  master seed (runmytrim.C:288)
//   scale L                   length scale
//   tmin T / cw C             SimconfType::tmin (default 0.2, as runmytrim) / SimconfType::cw (default 0.001)
//   tally vaccount|vacenergycount|range|base|primaries|recoils|phonon
//   primaries_only 0|1
//   potential universal|moliere|ckr   TrimBase::_potential (trim.h:63-69; default universal)
//   box wx wy wz              SampleLayers(wx, wy, wz); default wx = total thickness, 100, 100
//   sample layers|wire|burried_wire   sample class (default layers); wire samples take `box` as their size and
//                             the `layer` blocks as material[0], material[1] (thickness unused)
//   layer thickness rho nelem
//   elem Z m t [Edisp Elbind] (nelem lines after each layer)
//   start x y z dx dy dz      primary start (default 0, wy/2, wz/2, dir 1 0 0)
//   out prefix                writes prefix.records (binary mtb_record[N]) and prefix.hist
//   stopping Z m E            print getrstop of layer 0 for this ion (repeatable)
//   average Z m               print average() constants of layer 0
//   rng S k                   print k drand() (hex) after seeding with S, then k irand()
//   run
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <list>
#include <queue>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "simconf.h"
#include "element.h"
#include "material.h"
#include "sample_layers.h"
#include "sample_wire.h"
#include "sample_burried_wire.h"
#include "ion.h"
#include "trim.h"
#include "apps/include/ThreadedTrimBase.h"
#include "apps/include/TrimRange.h"
#include "apps/include/TrimVacCount.h"
#include "apps/include/TrimVacEnergyCount.h"

#include "../include/mytrim_b200.h"

using namespace MyTRIM_NS;

namespace
{
struct NullBuf : std::streambuf
{
  int overflow(int c) override { return c; }
};
NullBuf null_buf;
std::ostream null_stream(&null_buf);

struct Probe
{
  unsigned long steps = 0, replacements = 0, queued = 0;
};

// Adds per-cascade bookkeeping around any reference Trim class without changing what it does.
template <class Base>
struct Recording : Base
{
  template <class... A>
  Recording(A... a) : Base(a...)
  {
  }
  Probe probe;

protected:
  void replacementCollision() override
  {
    ++probe.replacements;
    Base::replacementCollision();
  }
  void checkPKAState() override
  {
    ++probe.steps;
    Base::checkPKAState();
  }
};

struct LayerDesc
{
  double thickness, rho;
  std::vector<Element> elements;
};

struct Worker
{
  SimconfType * simconf = nullptr;
  SampleBase * sample = nullptr;
  TrimBase * trim = nullptr;
  Probe * probe = nullptr;
  std::vector<unsigned> todo;
  unsigned long ions = 0;
};

struct Job
{
  int Z = 29;
  double m = 63.546, E = 1e4, Ef = 3.0;
  unsigned long n = 0;
  unsigned threads = 1;
  unsigned master = 2344;
  double scale = 1.0, tmin = 0.2, cw = -1.0;
  std::string tally = "vaccount", out, seedfile, potential = "universal", sample = "layers";
  bool primaries_only = false;
  bool have_box = false, have_start = false;
  double box[3] = {0, 100, 100};
  double start[6] = {0, 50, 50, 1, 0, 0};
  std::vector<LayerDesc> layers;
};

TrimBase *
makeTrim(const Job & job, SimconfType * sc, SampleBase * sample, Probe *& probe)
{
  if (job.tally == "vaccount")
  {
    auto * t = new Recording<TrimVacCount>(sc, sample);
    t->_primaries_only = job.primaries_only;
    probe = &t->probe;
    return t;
  }
  if (job.tally == "vacenergycount")
  {
    auto * t = new Recording<TrimVacEnergyCount>(sc, sample);
    t->_primaries_only = job.primaries_only;
    probe = &t->probe;
    return t;
  }
  if (job.tally == "range")
  {
    auto * t = new Recording<TrimRange>(sc, sample);
    t->_primaries_only = job.primaries_only;
    probe = &t->probe;
    return t;
  }
  if (job.tally == "primaries")
  {
    auto * t = new Recording<TrimPrimaries>(sc, sample);
    probe = &t->probe;
    return t;
  }
  if (job.tally == "recoils")
  {
    auto * t = new Recording<TrimRecoils>(sc, sample);
    probe = &t->probe;
    return t;
  }
  if (job.tally == "phonon")
  {
    // TrimPhononOut formats one text line per collision; a stream in the failed state makes every
    // operator<< return at its sentry, so only the EnucTotal bookkeeping remains
    null_stream.setstate(std::ios::badbit);
    auto * t = new Recording<TrimPhononOut>(sc, sample, std::ref(null_stream));
    probe = &t->probe;
    return t;
  }
  auto * t = new Recording<TrimBase>(sc, sample);
  probe = &t->probe;
  return t;
}

void
buildWorker(const Job & job, Worker & w)
{
  w.simconf = new SimconfType;
  w.simconf->fullTraj = false;
  w.simconf->tmin = job.tmin;
  if (job.cw > 0.0)
    w.simconf->cw = job.cw;
  w.simconf->setLengthScale(job.scale);
  double thickness = 0;
  for (auto & l : job.layers)
    thickness += l.thickness;
  SampleLayers * layered = nullptr;
  if (job.sample == "wire")
    w.sample = new SampleWire(job.box[0], job.box[1], job.box[2]);
  else if (job.sample == "burried_wire")
    w.sample = new SampleBurriedWire(job.box[0], job.box[1], job.box[2]);
  else if (job.have_box)
    w.sample = layered = new SampleLayers(job.box[0], job.box[1], job.box[2]);
  else
    w.sample = layered = new SampleLayers(thickness, 100.0, 100.0);
  w.trim = makeTrim(job, w.simconf, w.sample, w.probe);
  w.trim->_potential = job.potential == "moliere" ? TrimBase::MOLIERE : job.potential == "ckr" ? TrimBase::CKR : TrimBase::UNIVERSAL;
  for (auto & l : job.layers)
  {
    auto * mat = new MaterialBase(w.simconf, l.rho);
    for (auto & e : l.elements)
      mat->_element.push_back(e);
    mat->prepare();
    w.sample->material.push_back(mat);
    if (layered)
      layered->layerThickness.push_back(l.thickness);
  }
}

void
cascadeLoop(const Job * job, Worker * w, const std::vector<unsigned> * seeds, mtb_record * records)
{
  std::queue<IonBase *> recoils;
  for (unsigned idx : w->todo)
  {
    IonBase * pka = new IonBase(job->Z, job->m, job->E);
    pka->_Ef = job->Ef;
    pka->_gen = 0;
    pka->_pos = Point(job->start[0], job->start[1], job->start[2]);
    pka->_dir = Point(job->start[3], job->start[4], job->start[5]);
    pka->_seed = (*seeds)[idx];

    w->simconf->seed(pka->_seed);
    const int vac0 = w->simconf->vacancies_created;
    const double eel0 = w->simconf->EelTotal, enuc0 = w->simconf->EnucTotal;
    const Probe p0 = *w->probe;
    const unsigned long ions0 = w->ions;
    mtb_record rec;
    std::memset(&rec, 0, sizeof(rec));

    recoils.push(pka);
    bool is_primary = true; // FIFO: the first ion popped is the primary itself
    while (!recoils.empty())
    {
      IonBase * ion = recoils.front();
      recoils.pop();
      w->sample->averages(ion);
      const unsigned long s0 = w->probe->steps;
      w->trim->trim(ion, recoils);
      ++w->ions;
      if (is_primary)
      {
        is_primary = false;
        for (int i = 0; i < 3; ++i)
          rec.pos[i] = ion->_pos(i);
        rec.E = ion->_E;
        rec.state = ion->_state;
        rec.primary_steps = w->probe->steps - s0;
      }
      delete ion;
    }
    if (records)
    {
      rec.Eel = w->simconf->EelTotal - eel0;
      rec.Enuc = w->simconf->EnucTotal - enuc0;
      rec.vacancies = w->simconf->vacancies_created - vac0;
      rec.replacements = w->probe->replacements - p0.replacements;
      rec.steps = w->probe->steps - p0.steps;
      rec.ions = w->ions - ions0;
      records[idx] = rec;
    }
  }
}

int
runJob(const Job & job)
{
  std::vector<Worker> workers(job.threads);
  for (auto & w : workers)
    buildWorker(job, w);

  Job j = job;
  if (!j.have_start)
  {
    j.start[0] = 0.0;
    j.start[1] = workers[0].sample->w[1] / 2.0;
    j.start[2] = workers[0].sample->w[2] / 2.0;
  }

  std::vector<unsigned> seeds(j.n);
  if (!j.seedfile.empty())
  {
    std::ifstream sf(j.seedfile, std::ios::binary);
    sf.read(reinterpret_cast<char *>(seeds.data()), sizeof(unsigned) * j.n);
    if (!sf)
    {
      std::cerr << "cannot read seeds\n";
      return 1;
    }
  }
  else
  {
    workers[0].simconf->seed(j.master); // runmytrim.C:288-304
    for (auto & s : seeds)
      s = workers[0].simconf->irand();
  }
  for (unsigned long i = 0; i < j.n; ++i)
    workers[i % j.threads].todo.push_back(i);

  std::vector<mtb_record> records;
  if (!j.out.empty())
    records.resize(j.n);
  mtb_record * rp = records.empty() ? nullptr : records.data();

  const auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (auto & w : workers)
    th.emplace_back(cascadeLoop, &j, &w, &seeds, rp);
  for (auto & t : th)
    t.join();
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

  long vac = 0;
  double eel = 0, enuc = 0;
  unsigned long steps = 0, ions = 0, repl = 0;
  for (auto & w : workers)
  {
    vac += w.simconf->vacancies_created;
    eel += w.simconf->EelTotal;
    enuc += w.simconf->EnucTotal;
    steps += w.probe->steps;
    repl += w.probe->replacements;
    ions += w.ions;
  }
  if (!j.out.empty())
  {
    std::ofstream rf(j.out + ".records", std::ios::binary);
    rf.write(reinterpret_cast<const char *>(records.data()), sizeof(mtb_record) * records.size());
    // histograms through the reference's own threadJoin + writeOutput (runmytrim.C:316-326)
    auto * t0p = dynamic_cast<ThreadedTrimBase *>(workers[0].trim);
    if (t0p)
    {
      for (unsigned i = 1; i < j.threads; ++i)
        t0p->threadJoin(*dynamic_cast<ThreadedTrimBase *>(workers[i].trim));
      t0p->setBaseName(j.out);
      t0p->writeOutput();
    }
  }
  std::printf("{\"n\": %lu, \"threads\": %u, \"seconds\": %.6f, \"cascades_per_s\": %.3f, \"steps\": %lu, "
              "\"ions\": %lu, \"vacancies\": %ld, \"replacements\": %lu, \"Eel\": %.17g, \"Enuc\": %.17g}\n",
              j.n, j.threads, secs, j.n / secs, steps, ions, vac, repl, eel, enuc);
  return 0;
}
} // namespace

int
main()
{
  Job job;
  std::string line;
  int pending_elems = 0;
  while (std::getline(std::cin, line))
  {
    std::istringstream is(line);
    std::string cmd;
    if (!(is >> cmd) || cmd[0] == '#')
      continue;
    if (cmd == "ion")
    {
      is >> job.Z >> job.m >> job.E;
      if (!(is >> job.Ef))
        job.Ef = 3.0;
    }
    else if (cmd == "n")
      is >> job.n;
    else if (cmd == "threads")
      is >> job.threads;
    else if (cmd == "seeds")
      is >> job.seedfile;
    else if (cmd == "master")
      is >> job.master;
    else if (cmd == "scale")
      is >> job.scale;
    else if (cmd == "tmin")
      is >> job.tmin;
    else if (cmd == "cw")
      is >> job.cw;
    else if (cmd == "tally")
      is >> job.tally;
    else if (cmd == "primaries_only")
      is >> job.primaries_only;
    else if (cmd == "potential")
      is >> job.potential;
    else if (cmd == "sample")
      is >> job.sample;
    else if (cmd == "out")
      is >> job.out;
    else if (cmd == "box")
    {
      is >> job.box[0] >> job.box[1] >> job.box[2];
      job.have_box = true;
    }
    else if (cmd == "start")
    {
      for (double & v : job.start)
        is >> v;
      job.have_start = true;
    }
    else if (cmd == "layer")
    {
      LayerDesc l;
      is >> l.thickness >> l.rho >> pending_elems;
      job.layers.push_back(l);
    }
    else if (cmd == "elem")
    {
      Element e;
      is >> e._Z >> e._m >> e._t;
      double a, b;
      if (is >> a >> b)
      {
        e._Edisp = a;
        e._Elbind = b;
      }
      if (job.layers.empty() || pending_elems <= 0)
      {
        std::cerr << "elem without layer\n";
        return 1;
      }
      job.layers.back().elements.push_back(e);
      --pending_elems;
    }
    else if (cmd == "stopping" || cmd == "average")
    {
      Worker w;
      Job j1 = job;
      j1.tally = "base";
      buildWorker(j1, w);
      int Z;
      double m, E = 0;
      is >> Z >> m;
      if (cmd == "stopping")
        is >> E;
      IonBase ion(Z, m, E);
      MaterialBase * mat = w.sample->material[0];
      mat->average(&ion);
      if (cmd == "stopping")
        std::printf("stopping %d %.17g %.17g %.17g\n", Z, m, E, mat->getrstop(&ion));
      else
      {
        std::printf("average %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g", Z, m, mat->_arho, mat->_am,
                    mat->_az, mat->a, mat->f, mat->epsdg);
        for (auto & e : mat->_element)
          std::printf(" %.17g %.17g %.17g %.17g", e.my, e.ec, e.ai, e.fi);
        std::printf("\n");
      }
    }
    else if (cmd == "rng")
    {
      unsigned s, k;
      is >> s >> k;
      SimconfType sc;
      sc.seed(s);
      for (unsigned i = 0; i < k; ++i)
        std::printf("drand %a\n", sc.drand());
      for (unsigned i = 0; i < k; ++i)
        std::printf("irand %u\n", sc.irand());
    }
    else if (cmd == "run")
    {
      if (int rc = runJob(job))
        return rc;
    }
    else
    {
      std::cerr << "unknown command " << cmd << "\n";
      return 1;
    }
  }
  return 0;
}
