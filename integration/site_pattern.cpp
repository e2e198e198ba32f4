// integration/site_pattern.cpp -- replacement body of phylovi/libsbn's
// src/site_pattern.cpp (the header, src/site_pattern.hpp, is the reference's own):
// SitePattern::Compress, which turns the alignment into the (patterns, weights) pair
// every likelihood engine is built from, runs on the device through
// sbnb_compress_site_patterns (include/sbn_b200_patterns.h) instead of hashing
// std::vector<int> columns into an unordered_map on one host core
// (reference src/site_pattern.cpp:77-115).
//
// Same SET of (pattern, weight) pairs as the reference; the order is the order of first
// appearance in the alignment (the reference's is libstdc++'s hash-iteration order), which
// no likelihood depends on.

#include "site_pattern.hpp"

#include <algorithm>
#include <string>
#include <vector>

#include "intpack.hpp"
#include "sbn_b200_patterns.h"

// DNA: A C G T in either case are states 0..3; gap, unknown and every degenerate
// nucleotide code is state 4 (reference site_pattern.cpp:15-45, issue #162).
CharIntMap SitePattern::GetSymbolTable() {
  CharIntMap table;
  const std::string states = "ACGT";
  for (size_t state = 0; state < states.size(); state++) {
    table[states[state]] = static_cast<int>(state);
    table[static_cast<char>(states[state] - 'A' + 'a')] = static_cast<int>(state);
  }
  for (const char unresolved : std::string("-NX?BDHKMRSUVWY")) {
    table[unresolved] = 4;
  }
  return table;
}

int SitePattern::SymbolTableAt(const CharIntMap &symbol_table, char c) {
  const auto found = symbol_table.find(c);
  if (found == symbol_table.end()) {
    Failwith(std::string("Symbol '") + c + "' not known.");
  }
  return found->second;
}

SymbolVector SitePattern::SymbolVectorOf(const CharIntMap &symbol_table,
                                         const std::string &str) {
  SymbolVector symbols;
  symbols.reserve(str.size());
  for (const char c : str) {
    symbols.push_back(SymbolTableAt(symbol_table, c));
  }
  return symbols;
}

void SitePattern::Compress() {
  const size_t taxon_count = alignment_.SequenceCount();
  const size_t site_count = alignment_.Length();
  // Row t of the character matrix = the sequence of the taxon with leaf id t.
  std::string characters(taxon_count * site_count, '-');
  std::vector<bool> seen(taxon_count, false);
  for (const auto &[tag, taxon] : tag_taxon_map_) {
    const auto taxon_number = static_cast<size_t>(MaxLeafIDOfTag(tag));
    if (taxon_number >= taxon_count || seen[taxon_number]) {
      Failwith("SitePattern: taxon numbers must be distinct and below the sequence count.");
    }
    seen[taxon_number] = true;
    const std::string &sequence = alignment_.at(taxon);
    std::copy(sequence.begin(), sequence.end(), characters.begin() + taxon_number * site_count);
  }
  std::vector<uint8_t> patterns(taxon_count * site_count);
  weights_.assign(site_count, 0.);
  int64_t pattern_count = 0;
  if (sbnb_compress_site_patterns(static_cast<int32_t>(taxon_count),
                                  static_cast<int64_t>(site_count), characters.data(),
                                  /*device=*/0, patterns.data(), weights_.data(), &pattern_count,
                                  nullptr) != SBNB_OK) {
    Failwith(std::string("libsbn_b200: ") + sbnb_last_error());
  }
  weights_.resize(static_cast<size_t>(pattern_count));
  for (size_t taxon_number = 0; taxon_number < taxon_count; taxon_number++) {
    const uint8_t *row = patterns.data() + taxon_number * static_cast<size_t>(pattern_count);
    patterns_[taxon_number].assign(row, row + pattern_count);
  }
}

// Tip partials of one sequence, pattern-major: a one-hot row for a resolved state, all
// ones for state 4 (reference site_pattern.cpp:117-132; only BEAGLE's partials path used it).
const std::vector<double> SitePattern::GetPartials(size_t sequence_idx) const {
  const size_t state_count = 4;
  const SymbolVector &symbols = patterns_.at(sequence_idx);
  std::vector<double> partials(state_count * symbols.size(), 0.);
  for (size_t pattern = 0; pattern < symbols.size(); pattern++) {
    double *row = partials.data() + pattern * state_count;
    if (symbols[pattern] >= 0 && static_cast<size_t>(symbols[pattern]) < state_count) {
      row[symbols[pattern]] = 1.;
    } else {
      std::fill(row, row + state_count, 1.);
    }
  }
  return partials;
}
