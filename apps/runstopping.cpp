// runstopping — electronic stopping tables, same input as the reference's apps/runstopping.C:
//   ./runstopping < input.json      prints "E getrstop(E)" per energy [eV, eV/Ang]
// MaterialBase::getrstop runs the device function the transport kernel uses.
#include <cstdlib>
#include <iostream>
#include <vector>

#include "mini_json.h"
#include "mytrim/simconf.h"
#include "mytrim/material.h"
#include "mytrim/ion.h"

using namespace MyTRIM_NS;
using mini_json::Value;

static int
die(const std::string & msg)
{
  std::cerr << "ERROR: " << msg << '\n';
  return 1;
}

int
main(int argc, char **)
{
  if (argc > 1)
    return die("Please supply the input file via stdin (e.g. ./runstopping < input.json`)");
  Value root;
  try
  {
    root = Value::parse(std::cin);
  }
  catch (const std::exception & e)
  {
    return die(e.what());
  }
  if (!root["stopping"].isObject())
    return die("No 'stopping' top level block found in input");
  const Value & in = root["stopping"];

  SimconfType simconf;
  simconf.fullTraj = false;
  simconf.tmin = 0.2;

  if (!in["material"].isObject())
    return die("Must specify a 'material' block in the input file");
  if (!in["material"]["rho"].isNumeric())
    return die("Missing 'rho'");
  if (!in["material"]["elements"].isArray())
    return die("Missing 'elements' in material");
  MaterialBase material(&simconf, in["material"]["rho"].asDouble());
  const Value & els = in["material"]["elements"];
  for (size_t j = 0; j < els.size(); ++j)
  {
    Element element;
    for (const char * key : {"Z", "mass", "fraction"})
      if (!els[j][key].isNumeric())
        return die(std::string("Missing '") + key + "' in element " + std::to_string(j));
    element._Z = els[j]["Z"].asInt();
    element._m = els[j]["mass"].asDouble();
    element._t = els[j]["fraction"].asDouble();
    material._element.push_back(element);
  }
  material.prepare();

  if (!in["ion"].isObject())
    return die("Must specify an 'ion' block in the input file");
  if (!in["ion"]["Z"].isNumeric())
    return die("Missing 'Z' in ion block");
  if (!in["ion"]["mass"].isNumeric())
    return die("Missing 'mass' in ion block");
  IonBase pka(in["ion"]["Z"].asInt(), in["ion"]["mass"].asDouble(), 0.0);

  // single number, list, or {begin, end, step | mult}
  std::vector<Real> energies;
  const Value & en = in["ion"]["energy"];
  if (en.isNumeric())
    energies.push_back(en.asDouble());
  else if (en.isArray())
    for (size_t j = 0; j < en.size(); ++j)
      energies.push_back(en[j].asDouble());
  else if (en.isObject())
  {
    if (!en["begin"].isNumeric())
      return die("Missing 'begin' in energy block");
    if (!en["end"].isNumeric())
      return die("Missing 'end' in energy block");
    const bool step = en["step"].isNumeric(), mult = en["mult"].isNumeric();
    if (step == mult)
      return die("Specify either 'step' or 'mult' energy block");
    if (mult && en["mult"].asDouble() <= 1.0)
      return die("'mult' must be larger than 1.0");
    for (Real E = en["begin"].asDouble(); E <= en["end"].asDouble();
         E = step ? E + en["step"].asDouble() : E * en["mult"].asDouble())
      energies.push_back(E);
  }
  else
    return die("Missing or invalid 'energy' in ion block");

  for (Real E : energies)
  {
    pka._E = E;
    std::cout << pka._E << ' ' << material.getrstop(&pka) << '\n';
  }
  return EXIT_SUCCESS;
}
