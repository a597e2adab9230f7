// mytrim_uo2 — fission-fragment cascades in UO2 with Xe bubbles, same command line and output files as
// the reference's apps/mytrim_uo2.C:
//   ./mytrim_uo2 basename r Cbfactor Nev        (MYTRIM_SEED=<n> for a reproducible run)
// writes basename.clcoor (bubble coordinates), basename.Erec (energy, generation, MD tag of every Xe
// recoil) and basename.dist (displacement of every Xe recoil from the centre of its bubble of origin).
//
// Difference from the reference driver: all 2*Nev fragments are generated first (host mt19937, same
// inverse-CDF sampling) and followed in ONE batch on the GPU; the per-ion pre/post analysis of
// mytrim_uo2.C:281-338 runs afterwards on the engine's ion log.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
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

  // The ion log holds a birth and a death entry per Xe ion; events are processed in chunks so that
  // it stays bounded however many events are requested.
  const int chunk_events = std::getenv("MYTRIM_UO2_CHUNK") ? std::max(1, std::atoi(std::getenv("MYTRIM_UO2_CHUNK"))) : 32768;
  TrimXeLog trim(simconf, sample, 1ull << 22);
  MassInverter mass;
  EnergyInverter energy;
  Real Efiss = 0.0;
  std::vector<mtb_ion_log> log;
  double kernel_ms = 0.0; // device time of the transport launches (reported when MYTRIM_TIMING is set)
  unsigned long long n_primaries = 0;
  for (int first = 0; first < Nev; first += chunk_events)
  {
    // fission fragment pairs (mytrim_uo2.C:226-266)
    std::vector<IonBase *> primaries;
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
      primaries.push_back(ff1);
      primaries.push_back(ff2);
      Efiss += ff1->_E + ff2->_E;
    }

    if (!trim.trimBatch(primaries))
    {
      std::cerr << "ERROR: " << trim.lastError() << std::endl;
      return 1;
    }
    {
      float ms = 0.f;
      if (mtb_last_kernel_ms(trim.engine(), &ms) == MTB_OK)
        kernel_ms += ms;
      n_primaries += primaries.size();
    }
    for (auto * p : primaries)
      delete p;

    size_t n = 0;
    mtb_get_ion_log(trim.engine(), nullptr, 0, &n);
    log.resize(n);
    if (n && mtb_get_ion_log(trim.engine(), log.data(), n, &n) != MTB_OK)
    {
      std::cerr << "ERROR: " << mtb_last_error() << std::endl;
      return 1;
    }
    mtb_clear_lists(trim.engine());

    for (const auto & l : log)
    {
      // mark ions born in the MD energy gap (mytrim_uo2.C:285-287)
      const int md = (l.E0 > 200 && l.E0 < 12000) ? 1 : 0;
      if (l.gen > 0)
        std::fprintf(erec, "%f\t%d\t%d\n", l.E0, l.gen, md);
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
        std::fprintf(rdist, "%f %d %f %f %f\n", std::sqrt(d2), md, l.pos1[0], l.pos1[1], l.pos1[2]);
      }
    }
  }
  std::fclose(erec);
  std::fclose(rdist);

  if (std::getenv("MYTRIM_TIMING"))
  {
    mtb_counters cnt;
    if (mtb_get_counters(trim.engine(), &cnt) == MTB_OK && kernel_ms > 0.0)
      std::fprintf(stderr,
                   "{\"workload\": \"uo2_fission\", \"primaries\": %llu, \"collision_steps\": %llu, \"ions\": %llu, "
                   "\"kernel_ms\": %.3f, \"primaries_per_s\": %.4g, \"collision_steps_per_s\": %.4g}\n",
                   n_primaries, (unsigned long long)cnt.steps, (unsigned long long)cnt.ions, kernel_ms,
                   n_primaries / (kernel_ms * 1e-3), cnt.steps / (kernel_ms * 1e-3));
  }

  // energy accounting of the whole run (the reference prints it per event, mytrim_uo2.C:345-349)
  std::cout << simconf->EelTotal << std::endl;
  std::cout << simconf->EnucTotal << std::endl;
  std::cout << Efiss - (simconf->EelTotal + simconf->EnucTotal) << std::endl;
  return EXIT_SUCCESS;
}
