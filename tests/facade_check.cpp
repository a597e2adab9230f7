// facade_check.cpp — exercises the C++ plugin surface (include/mytrim) the way reference apps do;
// prints one JSON object that tests/test_gpu_facade.py compares with the oracle.  GPU only.
#include <cstdio>
#include <queue>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "mytrim/simconf.h"
#include "mytrim/sample_layers.h"
#include "mytrim/sample_solid.h"
#include "mytrim/trim.h"
#include "mytrim/include/TrimVacCount.h"

using namespace MyTRIM_NS;

// a user-defined Trim subclass with host-only hooks, as in the reference's documentation
class CountingTrim : public TrimBase
{
public:
  CountingTrim(SimconfType * s, SampleBase * b) : TrimBase(s, b), vac(0), repl(0), sub(0), followed(0), steps(0) {}
  long vac, repl, sub, followed, steps;
  double recoil_energy_sum = 0.0;

protected:
  virtual bool followRecoil()
  {
    ++followed;
    recoil_energy_sum += _recoil->_E;
    return true;
  }
  virtual void vacancyCreation()
  {
    ++vac;
    _simconf->vacancies_created++;
  }
  virtual void replacementCollision() { ++repl; }
  virtual void dissipateRecoilEnergy() { ++sub; }
  virtual void checkPKAState() { ++steps; }
};

static MaterialBase *
copper(SimconfType * sc)
{
  MaterialBase * m = new MaterialBase(sc, 8.92);
  Element e;
  e._Z = 29;
  e._m = 63.546;
  e._t = 1.0;
  m->_element.push_back(e);
  m->prepare();
  return m;
}

int
main()
{
  // --- 1. batched cascades with an in-tree tally class (what runmytrim does) ---
  SimconfType sc;
  sc.seed(2344);
  SampleLayers sample(1000.0, 100.0, 100.0);
  sample.material.push_back(copper(&sc));
  sample.layerThickness.push_back(1000.0);
  TrimVacCount trim(&sc, &sample);
  const int n = 2000;
  std::vector<IonBase *> prim;
  for (int i = 0; i < n; ++i)
  {
    IonBase * p = new IonBase(29, 63.546, 1.0e4);
    p->_gen = 0;
    p->_dir = Point(1, 0, 0);
    p->_pos = Point(0, 50, 50);
    prim.push_back(p);
  }
  std::vector<mtb_record> rec;
  if (!trim.trimBatch(prim, &rec))
  {
    std::fprintf(stderr, "trimBatch failed: %s\n", trim.lastError().c_str());
    return 1;
  }
  unsigned long hist_vac = 0, hist_repl = 0;
  for (unsigned v : trim.vacancies())
    hist_vac += v;
  for (unsigned v : trim.replacements())
    hist_repl += v;
  double xsum = 0;
  for (auto * p : prim)
    xsum += p->_pos(0);
  std::printf("{\"batch\": {\"n\": %d, \"vacancies\": %d, \"Eel\": %.10g, \"hist_vac\": %lu, \"hist_repl\": %lu, "
              "\"mean_x\": %.10g, \"bins\": %zu, \"rec0_vac\": %u, \"rec0_x\": %.10g},\n",
              n, sc.vacancies_created, sc.EelTotal, hist_vac, hist_repl, xsum / n, trim.vacancies().size(),
              rec[0].vacancies, rec[0].pos[0]);

  // --- 1b. trim() and trimBatch() alternating on the SAME object: the batch engine and its merged tallies survive a
  // single-ion call (two handles), and a configuration change between calls rebuilds the engine ---
  {
    std::queue<IonBase *> q;
    IonBase * one = new IonBase(29, 63.546, 1.0e4);
    one->_gen = 0;
    one->_dir = Point(1, 0, 0);
    one->_pos = Point(0, 50, 50);
    sample.averages(one);
    const int vac_before_single = sc.vacancies_created;
    trim.trim(one, q);
    const int single_vac = sc.vacancies_created - vac_before_single; // host hook of TrimVacCount ran
    while (!q.empty())
    {
      delete q.front();
      q.pop();
    }
    delete one;
    std::vector<IonBase *> prim2;
    for (int i = 0; i < n; ++i)
    {
      IonBase * p = new IonBase(29, 63.546, 1.0e4);
      p->_gen = 0;
      p->_dir = Point(1, 0, 0);
      p->_pos = Point(0, 50, 50);
      prim2.push_back(p);
    }
    if (!trim.trimBatch(prim2))
    {
      std::fprintf(stderr, "second trimBatch failed: %s\n", trim.lastError().c_str());
      return 1;
    }
    unsigned long hv2 = 0;
    for (unsigned v : trim.vacancies())
      hv2 += v;
    // a changed option must reach the device: tmin 0.2 -> 5 lengthens the free flights, fewer collisions
    const int vac_mid = sc.vacancies_created;
    sc.tmin = 5.0;
    std::vector<IonBase *> prim3;
    for (int i = 0; i < n; ++i)
    {
      IonBase * p = new IonBase(29, 63.546, 1.0e4);
      p->_gen = 0;
      p->_dir = Point(1, 0, 0);
      p->_pos = Point(0, 50, 50);
      prim3.push_back(p);
    }
    if (!trim.trimBatch(prim3))
    {
      std::fprintf(stderr, "third trimBatch failed: %s\n", trim.lastError().c_str());
      return 1;
    }
    unsigned long hv3 = 0;
    for (unsigned v : trim.vacancies())
      hv3 += v;
    sc.tmin = 0.2;
    std::printf(" \"alternate\": {\"single_vac\": %d, \"hist_vac_after_second\": %lu, \"vacancies_after_second\": %d, "
                "\"hist_vac_after_third\": %lu, \"vacancies_third\": %d},\n",
                single_vac, hv2, vac_mid, hv3, sc.vacancies_created - vac_mid);
  }

  // --- 2. the reference's per-ion loop with a user subclass: trim() + host hooks ---
  SimconfType sc2;
  sc2.seed(77);
  SampleSolid solid(1000.0, 100.0, 100.0);
  solid.material.push_back(copper(&sc2));
  CountingTrim ct(&sc2, &solid);
  const int n2 = 60;
  std::queue<IonBase *> recoils;
  long ions = 0;
  double primary_x = 0;
  for (int i = 0; i < n2; ++i)
  {
    IonBase * pka = new IonBase(29, 63.546, 1.0e4);
    pka->_gen = 0;
    pka->_dir = Point(1, 0, 0);
    pka->_pos = Point(0, 50, 50);
    recoils.push(pka);
    bool first = true;
    while (!recoils.empty())
    {
      IonBase * ion = recoils.front();
      recoils.pop();
      solid.averages(ion);
      ct.trim(ion, recoils);
      ++ions;
      if (first)
        primary_x += ion->_pos(0);
      first = false;
      delete ion;
    }
  }
  // --- 3. TrimDefectLog, TrimHistory and SimconfType::fullTraj through the reference's queue loop (trim.h:139-175,
  // trim.C:370-371, 421-422, 466-481): the hooks see every collision of every ion in the reference's order ---
  {
    SimconfType sc3;
    sc3.seed(4242);
    SampleSolid solid3(1000.0, 100.0, 100.0);
    solid3.material.push_back(copper(&sc3));
    std::ostringstream defects;
    TrimDefectLog dl(&sc3, &solid3, defects);
    TrimHistory hist(&sc3, &solid3);
    const int n3 = 40;
    long v_lines = 0, i_lines = 0, r_lines = 0, s_lines = 0, ions3 = 0, hist_followed = 0, hist_ions = 0;
    for (int pass = 0; pass < 2; ++pass)
    {
      std::queue<IonBase *> q;
      for (int i = 0; i < n3; ++i)
      {
        IonBase * pka = new IonBase(29, 63.546, 1.0e4);
        pka->_gen = 0;
        pka->_dir = Point(1, 0, 0);
        pka->_pos = Point(0, 50, 50);
        q.push(pka);
        while (!q.empty())
        {
          IonBase * ion = q.front();
          q.pop();
          solid3.averages(ion);
          if (pass == 0)
          {
            dl.trim(ion, q);
            ++ions3;
          }
          else
          {
            hist.trim(ion, q);
            ++hist_ions;
          }
          delete ion;
        }
      }
    }
    hist_followed = (long)hist.getHistory().size();
    std::istringstream in(defects.str());
    std::string line;
    while (std::getline(in, line))
    {
      if (line.rfind("V ", 0) == 0) ++v_lines;
      else if (line.rfind("I ", 0) == 0) ++i_lines;
      else if (line.rfind("R ", 0) == 0) ++r_lines;
      else if (line.rfind("S ", 0) == 0) ++s_lines;
    }
    // fullTraj: one "spawn" line per followed recoil and one state line per collision on stdout (captured here)
    SimconfType sc4;
    sc4.seed(99);
    sc4.fullTraj = true;
    SampleSolid solid4(1000.0, 100.0, 100.0);
    solid4.material.push_back(copper(&sc4));
    CountingTrim ct4(&sc4, &solid4);
    std::ostringstream traj;
    std::streambuf * old = std::cout.rdbuf(traj.rdbuf());
    long ions4 = 0;
    {
      std::queue<IonBase *> q;
      for (int i = 0; i < 5; ++i)
      {
        IonBase * pka = new IonBase(29, 63.546, 1.0e4);
        pka->_gen = 0;
        pka->_dir = Point(1, 0, 0);
        pka->_pos = Point(0, 50, 50);
        q.push(pka);
        while (!q.empty())
        {
          IonBase * ion = q.front();
          q.pop();
          ct4.trim(ion, q);
          ++ions4;
          delete ion;
        }
      }
    }
    std::cout.rdbuf(old);
    long spawn_lines = 0, state_lines = 0;
    std::istringstream tin(traj.str());
    while (std::getline(tin, line))
      (line.rfind("spawn ", 0) == 0 ? spawn_lines : state_lines)++;
    std::printf(" \"hooks\": {\"n\": %d, \"V\": %ld, \"I\": %ld, \"R\": %ld, \"S\": %ld, \"ions\": %ld, \"history\": %ld, "
                "\"history_ions\": %ld, \"spawn_lines\": %ld, \"state_lines\": %ld, \"traj_followed\": %ld, "
                "\"traj_steps\": %ld, \"traj_ions\": %ld},\n",
                n3, v_lines, i_lines, r_lines, s_lines, ions3, hist_followed, hist_ions, spawn_lines, state_lines, ct4.followed,
                ct4.steps, ions4);
  }

  std::printf(" \"single\": {\"n\": %d, \"vac\": %ld, \"repl\": %ld, \"sub\": %ld, \"followed\": %ld, \"steps\": %ld, "
              "\"ions\": %ld, \"Eel\": %.10g, \"simconf_vac\": %d, \"mean_x\": %.10g, \"mean_recoil_E\": %.10g}}\n",
              n2, ct.vac, ct.repl, ct.sub, ct.followed, ct.steps, ions, sc2.EelTotal, sc2.vacancies_created,
              primary_x / n2, ct.recoil_energy_sum / ct.followed);
  return 0;
}
