// mytrim_layers — layered-sample recoil statistics, same stdin format and stdout summary as the
// reference's apps/mytrim_layers.C (10000 x 500 keV Xe, TrimRecoils: follow generation < 2):
//   ./mytrim_layers basename < inputs/samplelayers_zro2_multilayer.in
// The per-ion post-analysis of the reference (sum of squared displacements of all non-primary
// ions, mytrim_layers.C:174-178) is computed from the engine's ion log.
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <string>
#include <vector>

#include "mytrim/simconf.h"
#include "mytrim/element.h"
#include "mytrim/material.h"
#include "mytrim/sample_layers.h"
#include "mytrim/ion.h"
#include "mytrim/trim.h"

using namespace MyTRIM_NS;

namespace
{
// TrimRecoils with the ion log switched on so the displacement statistics come back from the GPU
class TrimRecoilsLogged : public TrimRecoils
{
public:
  TrimRecoilsLogged(SimconfType * simconf, SampleBase * sample) : TrimRecoils(simconf, sample) {}

protected:
  virtual void deviceHooks(DeviceHooks & h) const
  {
    TrimRecoils::deviceHooks(h);
    h.tally_mask |= MTB_TALLY_IONLOG;
  }
};
} // namespace

int
main(int argc, char * argv[])
{
  if (argc != 2)
  {
    std::cerr << "syntax:\n" << argv[0] << " basename" << std::endl;
    return 1;
  }
  unsigned int seed = 0;
  if (const char * env = std::getenv("MYTRIM_SEED"))
    seed = (unsigned int)std::atoi(env);
  else
  {
    FILE * urand = std::fopen("/dev/urandom", "r");
    if (!urand || std::fread(&seed, sizeof(seed), 1, urand) != 1)
      return 1;
    std::fclose(urand);
  }
  SimconfType * simconf = new SimconfType(seed);
  simconf->fullTraj = false;
  simconf->tmin = 0.2;

  const auto restOfLine = []() { std::cin.ignore(std::numeric_limits<std::streamsize>::max(), '\n'); };
  Real sx, sy, sz;
  std::cin >> sx >> sy >> sz;
  restOfLine();
  std::cout << "SS " << sx << ' ' << sy << ' ' << sz << std::endl;

  int nmax = 10000;
  if (const char * env = std::getenv("MYTRIM_NPKA"))
    nmax = std::atoi(env);
  std::cout << "NN " << nmax << " PKAs" << std::endl;

  SampleLayers * sample = new SampleLayers(sx, sy, sz);
  TrimRecoilsLogged * trim = new TrimRecoilsLogged(simconf, sample);

  int nlayer;
  std::cin >> nlayer;
  restOfLine();
  std::cout << "n_layers=" << nlayer << std::endl;
  for (int i = 0; i < nlayer; ++i)
  {
    std::string name;
    Real thick, rho, nelem;
    std::cin >> name >> thick >> rho >> nelem;
    restOfLine();
    std::cout << "Layer: " << name << "  d=" << thick << "Ang  rho=" << rho << "g/ccm  n_elements=" << nelem << std::endl;
    MaterialBase * material = new MaterialBase(simconf, rho);
    for (int j = 0; j < nelem; ++j)
    {
      Element element;
      std::cin >> name >> element._Z >> element._m >> element._t;
      restOfLine();
      std::cout << "  Element: " << name << "  Z=" << element._Z << "  m=" << element._m << "  fraction=" << element._t
                << std::endl;
      material->_element.push_back(element);
    }
    material->prepare();
    sample->material.push_back(material);
    sample->layerThickness.push_back(thick);
  }

  const Real A = 131.0, E = 5.0e5;
  const int Z = 54; // 500 keV Xe
  std::vector<IonBase *> primaries;
  for (int n = 0; n < nmax; ++n)
  {
    IonBase * pka = new IonBase(Z, A, E);
    pka->_gen = 0;
    pka->_tag = -1;
    pka->_id = simconf->_id++;
    pka->_dir = Point(1, 0, 0);
    pka->_pos = Point(0, sample->w[1] / 2.0, sample->w[2] / 2.0);
    pka->setEf();
    primaries.push_back(pka);
  }
  if (!trim->trimBatch(primaries))
  {
    std::cerr << "ERROR: " << trim->lastError() << std::endl;
    return 1;
  }

  // displacement of every followed ion that is not a Xe projectile (mytrim_layers.C:174-186)
  size_t n = 0;
  mtb_get_ion_log(trim->engine(), nullptr, 0, &n);
  std::vector<mtb_ion_log> log(n);
  if (n && mtb_get_ion_log(trim->engine(), log.data(), n, &n) != MTB_OK)
  {
    std::cerr << "ERROR: " << mtb_last_error() << std::endl;
    return 1;
  }
  int nrec = 0;
  Real sum_r2 = 0.0;
  for (const auto & l : log)
  {
    if (l.Z != Z)
    {
      Real d2 = 0.0;
      for (int i = 0; i < 3; ++i)
        d2 += (l.pos0[i] - l.pos1[i]) * (l.pos0[i] - l.pos1[i]);
      sum_r2 += d2;
      ++nrec;
    }
    if (l.Z == 29)
      std::printf("RP %f %d %d\n", l.pos1[0], (int)l.primary, l.gen);
  }
  std::cout << "n=" << nrec << " sum_r2=" << sum_r2 << std::endl;
  return EXIT_SUCCESS;
}
