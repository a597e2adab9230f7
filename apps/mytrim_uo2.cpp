// mytrim_uo2 — fission-fragment cascades in UO2 with Xe bubbles, same command line and output files as
// the reference's apps/mytrim_uo2.C:
//   ./mytrim_uo2 basename r Cbfactor Nev        (MYTRIM_SEED=<n> for a reproducible run)
// writes basename.clcoor (bubble coordinates), basename.Erec (energy, generation, MD tag of every Xe
// recoil) and basename.dist (displacement of every Xe recoil from the centre of its bubble of origin).
//
// Difference from the reference driver: the fragments are generated in chunks of events (host mt19937, same
// inverse-CDF sampling) and every chunk is followed in ONE batch on a GPU; the per-ion pre/post analysis of
// mytrim_uo2.C:281-338 runs afterwards on the engine's ion log.  MYTRIM_GPUS=<n> deals the chunks over n GPUs
// (one host thread each); the output files are written in event order and do not depend on n.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <iostream>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "mytrim/simconf.h"
#include "mytrim/element.h"
#include "mytrim/material.h"
#include "mytrim/sample_clusters.h"
#include "mytrim/ion.h"
#include "mytrim/trim.h"
#include "mytrim/invert.h"

using namespace MyTRIM_NS;

namespace
{
const int kGasZ = 54;

// plain TrimBase (follow everything, count vacancies) + a log of every Xe ion
class TrimXeLog : public TrimBase
{
public:
  TrimXeLog(SimconfType * simconf, SampleBase * sample, unsigned long long capacity) : TrimBase(simconf, sample), _capacity(capacity) {}

protected:
  virtual void deviceHooks(DeviceHooks & h) const
  {
    h.known = true;
    h.tally_mask = MTB_TALLY_IONLOG;
    h.ionlog_z = kGasZ;
    h.ionlog_capacity = _capacity;
  }
  unsigned long long _capacity;
};
} // namespace

int
main(int argc, char * argv[])
{
  if (argc != 5)
  {
    std::cerr << "syntax:\n"
              << argv[0] << " basename r Cbfactor Nev\n\n"
              << "r Bubble radius in Ang\n"
              << "Cbfactor=1 => 7e-4 bubbles/nm^3\n"
              << "Nev  number of fission events (two fragemnts each)\n";
    return 1;
  }
  SimconfType * simconf = new SimconfType;
  int seed;
  if (const char * env = std::getenv("MYTRIM_SEED"))
    seed = std::atoi(env);
  else
  {
    FILE * urand = std::fopen("/dev/urandom", "r");
    if (!urand || std::fread(&seed, sizeof(int), 1, urand) != 1)
      return 1;
    std::fclose(urand);
  }
  simconf->seed(seed < 0 ? -seed : seed);

  const Real r = std::atof(argv[2]), Cbf = std::atof(argv[3]);
  const int Nev = std::atoi(argv[4]);

  sampleClusters * sample = new sampleClusters(400.0, 400.0, 400.0);
  sample->initSpatialhash(int(sample->w[0] / r) - 1, int(sample->w[1] / r) - 1, int(sample->w[2] / r) - 1);
  const Real v_sam = sample->w[0] * sample->w[1] * sample->w[2];
  const int n_cl = v_sam * 7.0e-7 * Cbf;
  std::cerr << "adding " << n_cl << " clusters...\n";
  sample->addRandomClusters(n_cl, r, 25.0, simconf);

  char fname[400];
  std::snprintf(fname, sizeof(fname), "%s.clcoor", argv[1]);
  FILE * ccf = std::fopen(fname, "wt");
  for (int i = 0; i < sample->cn; ++i)
    std::fprintf(ccf, "%f %f %f %f %d\n", sample->c[0][i], sample->c[1][i], sample->c[2][i], sample->c[3][i], i);
  std::fclose(ccf);
  std::cerr << "sample built.\n";

  // UO2 matrix and Xe bubbles (mytrim_uo2.C:163-184)
  Element element;
  MaterialBase * material = new MaterialBase(simconf, 10.0);
  element._Z = 92;
  element._m = 235.0;
  element._t = 1.0;
  material->_element.push_back(element);
  element._Z = 8;
  element._m = 16.0;
  element._t = 2.0;
  material->_element.push_back(element);
  material->prepare();
  sample->material.push_back(material);
  material = new MaterialBase(simconf, 3.5);
  element._Z = kGasZ;
  element._m = 132.0;
  element._t = 1.0;
  material->_element.push_back(element);
  material->prepare();
  sample->material.push_back(material);

  char ename[400], dname[400];
  std::snprintf(ename, sizeof(ename), "%s.Erec", argv[1]);
  std::snprintf(dname, sizeof(dname), "%s.dist", argv[1]);
  FILE * erec = std::fopen(ename, "wt");
  FILE * rdist = std::fopen(dname, "wt");

  // Events are processed in chunks (the ion log holds a birth and a death entry per Xe ion: it stays bounded
  // however many events are requested).  The chunks are dealt round-robin over the GPUs: the main thread draws
  // the fragments of chunk after chunk from the run's RNG (the source stream is sequential), one worker thread
  // per GPU follows them, and the output lines are written in event order — so the files do not depend on the
  // number of GPUs (Philox stream id of a fragment = its global index; the log is sorted by fragment and ion id).
  const int chunk_events = std::getenv("MYTRIM_UO2_CHUNK") ? std::max(1, std::atoi(std::getenv("MYTRIM_UO2_CHUNK"))) : 32768;
  int ngpu = std::getenv("MYTRIM_GPUS") ? std::max(1, std::atoi(std::getenv("MYTRIM_GPUS"))) : 1;
  ngpu = std::min(ngpu, std::max(1, mtb_device_count()));
  const int n_chunks = (Nev + chunk_events - 1) / chunk_events;
  // Two engines (own streams) per GPU by default: a launch of this workload ends in a tail in which a few CTAs finish
  // the last heavy cascades while most SMs idle (work is shared inside a CTA only); with a second engine the CTAs of
  // the next chunk's launch move onto the idle SMs.  Measured -5.5 % per chunk (profiles/r02_variant_sweeps.md).
  const int engines_per_gpu =
      std::getenv("MYTRIM_ENGINES_PER_GPU") ? std::max(1, std::min(4, std::atoi(std::getenv("MYTRIM_ENGINES_PER_GPU")))) : 2;
  const int nworkers = std::min(ngpu * engines_per_gpu, std::max(n_chunks, 1));

  struct Chunk
  {
    std::vector<IonBase *> primaries;
    uint64_t first_stream = 0;
    std::string erec, dist;
    bool done = false;
  };
  struct Device
  {
    std::unique_ptr<SimconfType> simconf;
    std::unique_ptr<TrimXeLog> trim;
    std::deque<int> todo;
    double kernel_ms = 0.0;
    double t_batch = 0.0, t_log = 0.0, t_format = 0.0, t_wait = 0.0; // host seconds per phase (MYTRIM_TIMING)
    std::string error;
  };
  std::vector<Chunk> chunks(n_chunks);
  std::vector<Device> devices(nworkers); // one engine each; engine d runs on GPU d % ngpu
  for (int d = 0; d < nworkers; ++d)
  {
    devices[d].simconf.reset(new SimconfType);
    devices[d].simconf->seed(seed < 0 ? -seed : seed); // same Philox key on every device
    devices[d].simconf->device = d % ngpu;
    // ion log: a birth and a death entry per Xe ion, ~16 entries per fragment in the gold geometry; 64 per fragment reserved
    devices[d].trim.reset(new TrimXeLog(devices[d].simconf.get(), sample,
                                        std::max<unsigned long long>(1ull << 22, 128ull * (unsigned long long)chunk_events)));
  }
  std::mutex mtx;
  std::condition_variable cv;
  bool no_more = false, failed = false;

  auto worker = [&](int d) {
    Device & dev = devices[d];
    std::vector<mtb_ion_log> log;
    char line[256];
    for (;;)
    {
      int c;
      auto now = [] { return std::chrono::steady_clock::now(); };
      auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double>(b - a).count();
      };
      const auto t0 = now();
      {
        std::unique_lock<std::mutex> lk(mtx);
        cv.wait(lk, [&] { return !dev.todo.empty() || no_more || failed; });
        if (failed || dev.todo.empty())
          return;
        c = dev.todo.front();
        dev.todo.pop_front();
      }
      Chunk & ch = chunks[c];
      const auto t1 = now();
      dev.t_wait += secs(t0, t1);
      dev.simconf->setStreamId(ch.first_stream);
      bool ok = dev.trim->trimBatch(ch.primaries);
      const auto t2 = now();
      dev.t_batch += secs(t1, t2);
      if (!ok)
        dev.error = dev.trim->lastError();
      size_t n = 0;
      if (ok)
      {
        float ms = 0.f;
        if (mtb_last_kernel_ms(dev.trim->engine(), &ms) == MTB_OK)
          dev.kernel_ms += ms;
        mtb_get_ion_log(dev.trim->engine(), nullptr, 0, &n);
        log.resize(n);
        if (n && mtb_get_ion_log(dev.trim->engine(), log.data(), n, &n) != MTB_OK)
        {
          ok = false;
          dev.error = std::string(mtb_last_error()) + " (fewer fission events per launch: MYTRIM_UO2_CHUNK=<events>)";
        }
        mtb_clear_lists(dev.trim->engine());
      }
      for (auto * p : ch.primaries)
        delete p;
      ch.primaries.clear();
      const auto t3 = now();
      dev.t_log += secs(t2, t3);
      if (ok)
      {
        // the device appends log entries in scheduling order: sort by fragment, then ion id
        std::sort(log.begin(), log.end(), [](const mtb_ion_log & a, const mtb_ion_log & b) {
          return a.primary != b.primary ? a.primary < b.primary : a.uid < b.uid;
        });
        for (const auto & l : log)
        {
          // mark ions born in the MD energy gap (mytrim_uo2.C:285-287)
          const int md = (l.E0 > 200 && l.E0 < 12000) ? 1 : 0;
          if (l.gen > 0)
          {
            std::snprintf(line, sizeof(line), "%f\t%d\t%d\n", l.E0, l.gen, md);
            ch.erec += line;
          }
          if (l.tag >= 0)
          {
            // displacement from the centre of the bubble of origin, minimum image (mytrim_uo2.C:296-337)
            Real d2 = 0.0;
            for (int i = 0; i < 3; ++i)
            {
              Real dif = sample->c[i][l.tag] - l.pos0[i];
              if (sample->bc[i] == SampleBase::PBC)
                dif -= std::round(dif / sample->w[i]) * sample->w[i];
              const Real centre = l.pos0[i] + dif;
              d2 += (centre - l.pos1[i]) * (centre - l.pos1[i]);
            }
            std::snprintf(line, sizeof(line), "%f %d %f %f %f\n", std::sqrt(d2), md, l.pos1[0], l.pos1[1], l.pos1[2]);
            ch.dist += line;
          }
        }
      }
      dev.t_format += secs(t3, now());
      {
        std::lock_guard<std::mutex> lk(mtx);
        ch.done = true;
        if (!ok)
          failed = true;
      }
      cv.notify_all();
    }
  };
  std::vector<std::thread> threads;
  for (int d = 0; d < nworkers; ++d)
    threads.emplace_back(worker, d);
  const auto t_transport0 = std::chrono::steady_clock::now();

  MassInverter mass;
  EnergyInverter energy;
  Real Efiss = 0.0;
  unsigned long long n_primaries = 0, n_degenerate = 0;
  int next_to_write = 0;
  auto flush_done = [&](std::unique_lock<std::mutex> &) {
    while (next_to_write < n_chunks && chunks[next_to_write].done)
    {
      Chunk & ch = chunks[next_to_write++];
      std::fputs(ch.erec.c_str(), erec);
      std::fputs(ch.dist.c_str(), rdist);
      std::string().swap(ch.erec);
      std::string().swap(ch.dist);
    }
  };
  for (int c = 0; c < n_chunks && !failed; ++c)
  {
    const int first = c * chunk_events;
    Chunk & ch = chunks[c];
    ch.first_stream = 2ull * (uint64_t)first;
    // fission fragment pairs (mytrim_uo2.C:226-266)
    for (int n = first; n < std::min(Nev, first + chunk_events); ++n)
    {
      const Real A1 = mass.x(simconf->drand());
      const Real A2 = 235.0 - A1;
      energy.setMass(A1);
      const Real Etot = energy.x(simconf->drand());
      const Real E1 = Etot * A2 / (A1 + A2), E2 = Etot - E1;
      const int Z1 = std::round((A1 * 92.0) / 235.0), Z2 = 92 - Z1;
      IonMDTag * ff1 = new IonMDTag;
      ff1->_gen = 0;
      ff1->_tag = -1;
      ff1->_Z = Z1;
      ff1->_m = A1;
      ff1->_E = E1 * 1.0e6;
      Real norm;
      do
      {
        for (int i = 0; i < 3; ++i)
          ff1->_dir(i) = 2.0 * simconf->drand() - 1.0;
        norm = ff1->_dir.norm_sq();
      } while (norm <= 0.0001 || norm > 1.0);
      ff1->_dir /= std::sqrt(norm);
      for (int i = 0; i < 3; ++i)
        ff1->_pos(i) = simconf->drand() * sample->w[i];
      ff1->setEf();
      IonMDTag * ff2 = new IonMDTag(*ff1);
      ff2->_dir = -ff2->_dir;
      ff2->_Z = Z2;
      ff2->_m = A2;
      ff2->_E = E2 * 1.0e6;
      ff2->setEf();
      for (IonMDTag * ff : {ff1, ff2})
        if (ff->_Z < 1 || ff->_Z > 92)
        {
          // About one draw in 1e6 ends the bisection of Inverter::x at A = 235 * 2^-33, i.e. Z = 0: the reference
          // then reads scoef[-1] (undefined behaviour).  The fragment keeps its place (its index is its Philox
          // stream) but carries no energy and stops where it starts.
          ff->_Z = ff->_Z < 1 ? 1 : 92;
          ff->_m = std::max(ff->_m, 1.0);
          ff->_E = 0.0;
          ++n_degenerate;
        }
      ch.primaries.push_back(ff1);
      ch.primaries.push_back(ff2);
      Efiss += ff1->_E + ff2->_E;
    }
    n_primaries += ch.primaries.size();
    {
      // at most two chunks in flight per GPU: the fragments of 1e8 primaries never exist at the same time
      std::unique_lock<std::mutex> lk(mtx);
      cv.wait(lk, [&] {
        flush_done(lk);
        return failed || c - next_to_write < 2 * nworkers;
      });
      devices[c % nworkers].todo.push_back(c);
    }
    cv.notify_all();
  }
  {
    std::unique_lock<std::mutex> lk(mtx);
    no_more = true;
    cv.notify_all();
    cv.wait(lk, [&] {
      flush_done(lk);
      return failed || next_to_write == n_chunks;
    });
  }
  cv.notify_all();
  for (auto & t : threads)
    t.join();
  const double wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_transport0).count();
  std::fclose(erec);
  std::fclose(rdist);
  if (failed)
  {
    for (auto & dev : devices)
      if (!dev.error.empty())
        std::cerr << "ERROR: " << dev.error << std::endl;
    return 1;
  }

  // join the per-GPU totals (the analogue of threadJoin)
  double kernel_ms = 0.0; // slowest engine: the devices run concurrently (engines of one GPU overlap: see wall_ms)
  unsigned long long steps = 0, ions = 0;
  for (auto & dev : devices)
  {
    simconf->EelTotal += dev.simconf->EelTotal;
    simconf->EnucTotal += dev.simconf->EnucTotal;
    simconf->vacancies_created += dev.simconf->vacancies_created;
    kernel_ms = std::max(kernel_ms, dev.kernel_ms);
    mtb_counters cnt;
    if (dev.trim->engine() && mtb_get_counters(dev.trim->engine(), &cnt) == MTB_OK)
    {
      steps += cnt.steps;
      ions += cnt.ions;
    }
  }
  if (std::getenv("MYTRIM_TIMING"))
    for (size_t d = 0; d < devices.size(); ++d)
      std::fprintf(stderr, "engine %zu: waiting for chunks %.2f s, trimBatch %.2f s (kernels %.2f s), ion log %.2f s, sort + format %.2f s\n", d,
                   devices[d].t_wait, devices[d].t_batch, devices[d].kernel_ms * 1e-3, devices[d].t_log, devices[d].t_format);
  if (std::getenv("MYTRIM_TIMING") && kernel_ms > 0.0)
    std::fprintf(stderr,
                 "{\"workload\": \"uo2_fission\", \"gpus\": %d, \"primaries\": %llu, \"collision_steps\": %llu, "
                 "\"ions\": %llu, \"kernel_ms\": %.3f, \"primaries_per_s\": %.4g, \"collision_steps_per_s\": %.4g, "
                 "\"engines_per_gpu\": %d, \"wall_ms\": %.3f, \"primaries_per_s_wall\": %.4g, "
                 "\"collision_steps_per_s_wall\": %.4g}\n",
                 ngpu, n_primaries, steps, ions, kernel_ms, n_primaries / (kernel_ms * 1e-3), steps / (kernel_ms * 1e-3),
                 (nworkers + ngpu - 1) / ngpu, wall_ms, n_primaries / (wall_ms * 1e-3), steps / (wall_ms * 1e-3));

  if (n_degenerate)
    std::cerr << "WARNING: " << n_degenerate << " fragment(s) with Z outside 1..92 were emitted without energy" << std::endl;
  // energy accounting of the whole run (the reference prints it per event, mytrim_uo2.C:345-349)
  std::cout << simconf->EelTotal << std::endl;
  std::cout << simconf->EnucTotal << std::endl;
  std::cout << Efiss - (simconf->EelTotal + simconf->EnucTotal) << std::endl;
  return EXIT_SUCCESS;
}
