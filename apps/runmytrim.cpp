// runmytrim — same input files and output files as the reference's apps/runmytrim.C, with the
// cascades running on B200 GPUs:  ./runmytrim < input.json
//
// Differences from the reference driver, all forced by where the work runs:
//   * options.threads is accepted and ignored; options.gpus (default 1) selects how many GPUs the
//     primaries are sharded over (contiguous index ranges, one host thread per GPU, tallies joined
//     with threadJoin exactly like the reference joins its threads).
//   * all primaries of a GPU go to the device in one TrimBase::trimBatch() call instead of the
//     per-ion pop/averages/trim loop (runmytrim.C:76-92).
//   * random numbers come from Philox streams keyed by (options.seed, primary index), so results do
//     not depend on the GPU count.
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "mini_json.h"
#include "mytrim/simconf.h"
#include "mytrim/element.h"
#include "mytrim/material.h"
#include "mytrim/sample_layers.h"
#include "mytrim/ion.h"
#include "mytrim/include/TrimRange.h"
#include "mytrim/include/TrimVacCount.h"
#include "mytrim/include/TrimVacEnergyCount.h"

using namespace MyTRIM_NS;
using mini_json::Value;

namespace
{
struct Shard
{
  std::unique_ptr<SimconfType> simconf;
  std::unique_ptr<SampleLayers> sample;
  std::unique_ptr<ThreadedTrimBase> trim;
  std::vector<IonBase *> primaries;
  uint64_t first = 0;
  bool ok = true;
  std::string error;
};

int
die(const std::string & msg)
{
  std::cerr << "ERROR: " << msg << '\n';
  return 1;
}
} // namespace

int
main(int argc, char **)
{
  if (argc > 1)
    return die("Please supply the input file via stdin (e.g. ./runmytrim < input.json`)");

  Value root;
  try
  {
    root = Value::parse(std::cin);
  }
  catch (const std::exception & e)
  {
    return die(e.what());
  }
  if (!root["mytrim"].isObject())
    return die("No 'mytrim' top level block found in input");
  const Value & in = root["mytrim"];
  const Value & opt = in["options"];

  unsigned int ngpu = 1;
  if (opt.isObject() && opt["gpus"].isNumeric())
    ngpu = std::max(1, opt["gpus"].asInt());
  if (opt.isObject() && opt["threads"].isNumeric())
    std::cerr << "Ignoring 'threads': cascades run on " << ngpu << " GPU(s)\n";

  long long master_seed = 0;
  if (opt.isObject() && opt["seed"].isNumeric())
  {
    master_seed = opt["seed"].asInt64();
    std::cerr << "Using provided master seed " << master_seed << '\n';
  }
  else
  {
    FILE * urand = std::fopen("/dev/urandom", "r");
    if (!urand || std::fread(&master_seed, sizeof(int), 1, urand) != 1)
      return die("Unable to access /dev/urandom");
    std::fclose(urand);
  }
  double scale = 1.0;
  if (opt.isObject() && opt["scale"].isNumeric())
  {
    scale = opt["scale"].asDouble();
    std::cerr << "Using provided length scale " << scale << '\n';
  }

  if (!in["sample"].isObject())
    return die("Must specify a 'sample' block in the input file");
  const Value & layers = in["sample"]["layers"];
  if (!layers.isArray())
    return die("sample.layers must be an array");
  std::cerr << "Building " << layers.size() << " layers\n";
  double thickness = 0.0;
  for (size_t i = 0; i < layers.size(); ++i)
  {
    if (!layers[i]["thickness"].isNumeric())
      return die("No 'thickness' found for layer " + std::to_string(i));
    thickness += layers[i]["thickness"].asDouble();
  }

  if (!in["output"].isObject())
    return die("Must specify an 'output' block in the input file");
  if (!in["output"]["type"].isString())
    return die("output.type must be a string");
  const std::string type = in["output"]["type"].asString();
  if (type != "vaccount" && type != "vacenergycount" && type != "range")
    return die("Unknown output type " + type);

  if (!in["ion"].isObject())
    return die("Must specify an 'ion' block in the input file");
  const Value & ion = in["ion"];
  for (const char * key : {"Z", "mass", "energy", "number"})
    if (!ion[key].isNumeric())
      return die(std::string("Missing '") + key + "' in ion block");
  const unsigned long npka = (unsigned long)ion["number"].asInt64();

  // one object set per GPU (the reference builds one per thread, runmytrim.C:60-68, 183-258)
  std::vector<Shard> shards(ngpu);
  for (unsigned int g = 0; g < ngpu; ++g)
  {
    Shard & s = shards[g];
    s.simconf.reset(new SimconfType);
    s.simconf->device = (int)g;
    s.simconf->fullTraj = false;
    s.simconf->tmin = 0.2;
    s.simconf->setLengthScale(scale);
    s.simconf->seed((unsigned int)master_seed);
    s.sample.reset(new SampleLayers(thickness, 100.0, 100.0));
    if (type == "vaccount")
      s.trim.reset(new TrimVacCount(s.simconf.get(), s.sample.get()));
    else if (type == "vacenergycount")
      s.trim.reset(new TrimVacEnergyCount(s.simconf.get(), s.sample.get()));
    else
      s.trim.reset(new TrimRange(s.simconf.get(), s.sample.get()));
    if (in["output"]["base"].isString())
      s.trim->setBaseName(in["output"]["base"].asString());
    if (in["output"]["primaries_only"].isBool())
      s.trim->_primaries_only = in["output"]["primaries_only"].asBool();

    for (size_t i = 0; i < layers.size(); ++i)
    {
      if (!layers[i]["rho"].isNumeric())
        return die("Missing 'rho' in layer " + std::to_string(i));
      if (!layers[i]["elements"].isArray())
        return die("Missing 'elements' in layer " + std::to_string(i));
      MaterialBase * material = new MaterialBase(s.simconf.get(), layers[i]["rho"].asDouble());
      const Value & els = layers[i]["elements"];
      for (size_t j = 0; j < els.size(); ++j)
      {
        Element element;
        for (const char * key : {"Z", "mass", "fraction"})
          if (!els[j][key].isNumeric())
            return die(std::string("Missing '") + key + "' in element " + std::to_string(j) + " in layer " +
                       std::to_string(i));
        element._Z = els[j]["Z"].asInt();
        element._m = els[j]["mass"].asDouble();
        element._t = els[j]["fraction"].asDouble();
        if (els[j]["edisp"].isNumeric())
          element._Edisp = els[j]["edisp"].asDouble();
        if (els[j]["elbind"].isNumeric())
          element._Elbind = els[j]["elbind"].asDouble();
        material->_element.push_back(element);
      }
      material->prepare();
      s.sample->material.push_back(material);
      s.sample->layerThickness.push_back(layers[i]["thickness"].asDouble());
    }
  }

  // primaries: contiguous index ranges per GPU (runmytrim.C:291-304 deals them round-robin to threads)
  IonBase proto(ion["Z"].asInt(), ion["mass"].asDouble(), ion["energy"].asDouble());
  if (ion["final_energy"].isNumeric())
    proto._Ef = ion["final_energy"].asDouble();
  for (unsigned int g = 0; g < ngpu; ++g)
  {
    Shard & s = shards[g];
    const unsigned long lo = npka * g / ngpu, hi = npka * (g + 1) / ngpu;
    s.first = lo;
    s.primaries.reserve(hi - lo);
    for (unsigned long n = lo; n < hi; ++n)
    {
      IonBase * pka = new IonBase(&proto);
      pka->_gen = 0;
      pka->_dir = Point(1.0, 0.0, 0.0);
      pka->_pos = Point(0.0, s.sample->w[1] / 2.0, s.sample->w[2] / 2.0);
      s.primaries.push_back(pka);
    }
  }

  std::vector<std::thread> workers;
  for (unsigned int g = 0; g < ngpu; ++g)
    workers.emplace_back([&shards, g]() {
      Shard & s = shards[g];
      s.simconf->nextStreamId(s.first); // global primary index == Philox stream id
      if (!s.trim->trimBatch(s.primaries))
      {
        s.ok = false;
        s.error = s.trim->lastError();
      }
    });
  for (auto & w : workers)
    w.join();
  for (auto & s : shards)
    if (!s.ok)
      return die(s.error);

  // join the per-GPU tallies into shard 0 (runmytrim.C:316-323)
  for (unsigned int g = 1; g < ngpu; ++g)
  {
    shards[0].trim->threadJoin(*shards[g].trim);
    shards[0].simconf->vacancies_created += shards[g].simconf->vacancies_created;
    shards[0].simconf->EelTotal += shards[g].simconf->EelTotal;
    shards[0].simconf->EnucTotal += shards[g].simconf->EnucTotal;
  }
  shards[0].trim->writeOutput();

  std::cerr << "Vacancies/ion: " << Real(shards[0].simconf->vacancies_created) / Real(npka) << '\n'
            << "Electronic energy loss/ion: " << shards[0].simconf->EelTotal / Real(npka) << '\n';
  for (auto & s : shards)
    for (auto * p : s.primaries)
      delete p;
  return EXIT_SUCCESS;
}
