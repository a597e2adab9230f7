// facade_host_check.cpp — the parts of the C++ plugin surface that are pure host code (no GPU needed):
// table loading, materials, samples' lookupMaterial, cluster placement, inverters, ion bookkeeping.
// Prints one JSON object that tests/test_facade_host.py compares with the oracle.
#include <cstdio>
#include <cstdlib>
#include <queue>
#include <vector>

#include "mytrim/simconf.h"
#include "mytrim/sample_layers.h"
#include "mytrim/sample_solid.h"
#include "mytrim/sample_wire.h"
#include "mytrim/sample_burried_wire.h"
#include "mytrim/sample_clusters.h"
#include "mytrim/invert.h"
#include "mytrim/trim.h"

using namespace MyTRIM_NS;

int
main()
{
  SimconfType sc(39172);
  std::printf("{\"drand\": [");
  for (int i = 0; i < 8; ++i)
    std::printf("%s%.17g", i ? ", " : "", sc.drand());
  std::printf("], \"irand\": [");
  for (int i = 0; i < 8; ++i)
    std::printf("%s%u", i ? ", " : "", sc.irand());
  std::printf("],\n \"scoef\": [");
  for (int z : {1, 8, 29, 54, 92})
    std::printf("%s[%.17g, %.17g, %.17g, %.17g, %.17g]", z == 1 ? "" : ", ", sc.scoef[z - 1].mm1, sc.scoef[z - 1].vfermi,
                sc.scoef[z - 1].lfctr, sc.scoef[z - 1].pcoef[0], sc.scoef[z - 1].pcoef[7]);
  std::printf("],\n");

  // UO2: prepare() + average() for a Xe ion
  MaterialBase uo2(&sc, 10.97);
  Element e;
  e._Z = 92;
  e._m = 238.03;
  e._t = 1.0;
  uo2._element.push_back(e);
  e._Z = 8;
  e._m = 15.999;
  e._t = 2.0;
  uo2._element.push_back(e);
  uo2.prepare();
  IonBase xe(54, 131.904, 1.0e7);
  uo2.average(&xe);
  std::printf(" \"average\": [%.17g, %.17g, %.17g, %.17g, %.17g, %.17g", uo2._arho, uo2._am, uo2._az, uo2.a, uo2.f, uo2.epsdg);
  for (auto & el : uo2._element)
    std::printf(", %.17g, %.17g, %.17g, %.17g", el.my, el.ec, el.ai, el.fi);
  std::printf("],\n");

  // cluster placement with the host RNG (mytrim_uo2.C:95-142) and lookups
  sc.seed(39172);
  sampleClusters cl(400.0, 400.0, 400.0);
  cl.initSpatialhash(39, 39, 39);
  cl.addRandomClusters(4, 10.0, 25.0, &sc);
  MaterialBase matrix(&sc, 10.0), gas(&sc, 3.5);
  cl.material.push_back(&matrix);
  cl.material.push_back(&gas);
  std::printf(" \"clusters\": [");
  for (int i = 0; i < cl.cn; ++i)
    std::printf("%s[%.17g, %.17g, %.17g, %.17g]", i ? ", " : "", cl.c[0][i], cl.c[1][i], cl.c[2][i], cl.c[3][i]);
  std::printf("],\n \"cluster_lookup\": [");
  for (int i = 0; i < 400; ++i)
  {
    // points around cluster 0 and wrapped through the periodic box
    Point p(cl.c[0][0] + 0.09 * (i % 20) * (i % 3 - 1) - 400.0 * (i % 2), cl.c[1][0] + 0.7 * (i / 20) - 6.0, cl.c[2][0] + 0.05 * i - 8.0);
    std::printf("%s%d", i ? ", " : "", cl.lookupCluster(p, 0.0));
  }
  std::printf("],\n");

  // layered / wire samples
  SampleLayers layers(30.0, 100.0, 100.0);
  MaterialBase m0(&sc, 1.0), m1(&sc, 2.0), m2(&sc, 3.0);
  for (auto * m : {&m0, &m1, &m2})
  {
    layers.material.push_back(m);
    layers.layerThickness.push_back(10.0);
  }
  std::printf(" \"layers\": [");
  const double xs[] = {-5.0, 0.0, 9.999, 10.0, 19.5, 20.0, 29.9, 30.0, 1e6};
  for (int i = 0; i < 9; ++i)
  {
    Point p(xs[i], 0, 0);
    std::printf("%s%d", i ? ", " : "", layers.lookupLayer(p));
  }
  SampleWire wire(200.0, 100.0, 50.0);
  wire.material.push_back(&m0);
  SampleBurriedWire bw(200.0, 100.0, 50.0);
  bw.material.push_back(&m0);
  bw.material.push_back(&m1);
  std::printf("],\n \"wire\": [");
  const double pts[][3] = {{100, 50, 10}, {1, 1, 10}, {199, 50, 10}, {100, 50, -100}, {100, 50, -251}, {1, 1, 25}, {100, 50, 51}};
  for (int i = 0; i < 7; ++i)
  {
    Point p(pts[i][0], pts[i][1], pts[i][2]);
    MaterialBase * a = wire.lookupMaterial(p);
    MaterialBase * b = bw.lookupMaterial(p);
    std::printf("%s[%d, %d]", i ? ", " : "", a ? 0 : -1, b == &m0 ? 0 : (b == &m1 ? 1 : -1));
  }
  std::printf("],\n \"bc\": [%d, %d, %d, %d],\n", (int)wire.bc[0], (int)wire.bc[2], (int)bw.bc[0], (int)layers.bc[0]);

  // fission yield inverters (invert.C)
  MassInverter mi;
  EnergyInverter ei;
  std::printf(" \"mass_x\": [");
  for (int i = 1; i < 10; ++i)
    std::printf("%s%.17g", i > 1 ? ", " : "", mi.x(0.1 * i));
  std::printf("], \"energy_x\": [");
  ei.setMass(96.0);
  for (int i = 1; i < 10; ++i)
    std::printf("%s%.17g", i > 1 ? ", " : "", ei.x(0.1 * i));
  std::printf("],\n");

  // ion bookkeeping (ion.C)
  IonMDTag parent;
  parent._gen = 3;
  parent._pos = Point(1, 2, 3);
  parent._Ef = 7.0;
  parent._md = 5;
  IonBase * child = parent.spawnRecoil();
  std::printf(" \"ion\": [%d, %.1f, %.1f, %d, %d, %d]}\n", child->_gen, child->_pos(2), child->_Ef, child->_tag,
              dynamic_cast<IonMDTag *>(child) ? dynamic_cast<IonMDTag *>(child)->_md : -1, (int)child->_state);
  delete child;
  return 0;
}
