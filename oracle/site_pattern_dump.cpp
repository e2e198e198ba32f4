// oracle/site_pattern_dump.cpp -- TEST INFRASTRUCTURE (fixture generator).
//
// Runs the UNMODIFIED reference SitePattern::Compress (src/site_pattern.cpp:77-115,
// objects compiled into oracle/_ref/obj by oracle/Makefile) on a FASTA file and
// prints, as JSON, the sequences in the order it numbered them and the patterns
// and weights it produced (in ITS order, the iteration order of an
// unordered_map).  tests/golden/make_site_pattern_fixtures.py turns that into
// the fixtures the device compression is pinned to.  Nothing of the reference is
// copied or modified.
//
// usage: site_pattern_dump <fasta>      taxon i = the i-th name in sorted order

#include <algorithm>
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>

#include "alignment.hpp"
#include "intpack.hpp"
#include "site_pattern.hpp"

int main(int argc, char** argv) {
  if (argc != 2) {
    std::fprintf(stderr, "usage: site_pattern_dump <fasta>\n");
    return 2;
  }
  const Alignment alignment = Alignment::ReadFasta(argv[1]);
  std::vector<std::string> names;
  for (const auto& [name, sequence] : alignment.Data()) names.push_back(name);
  std::sort(names.begin(), names.end());
  TagStringMap tag_taxon_map;
  for (size_t i = 0; i < names.size(); i++) tag_taxon_map[PackInts(static_cast<uint32_t>(i), 1)] = names[i];
  const SitePattern site_pattern(alignment, tag_taxon_map);
  std::cout << "{\"sequences\": [";
  for (size_t i = 0; i < names.size(); i++)
    std::cout << (i ? ", " : "") << "\"" << alignment.Data().at(names[i]) << "\"";
  std::cout << "], \"patterns\": [";
  const auto& patterns = site_pattern.GetPatterns();
  for (size_t t = 0; t < patterns.size(); t++) {
    std::cout << (t ? ", " : "") << "[";
    for (size_t k = 0; k < patterns[t].size(); k++) std::cout << (k ? "," : "") << patterns[t][k];
    std::cout << "]";
  }
  std::cout << "], \"weights\": [";
  const auto& weights = site_pattern.GetWeights();
  for (size_t k = 0; k < weights.size(); k++) std::cout << (k ? "," : "") << weights[k];
  std::cout << "]}\n";
  return 0;
}
