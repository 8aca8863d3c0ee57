// process_seq.hpp -- C++ host mirror of the reference's sequencing interface, above the C ABI.
//
// Same names, argument order, defaults and error behaviour as
//   simulate_seq()          src/seq_simulation.hpp:28-41,  src/seq_simulation.cpp:517-601
//   simulate_normal_seq()   src/seq_simulation.hpp:43-54,  src/seq_simulation.cpp:603-679
//   BasicIlluminaSequencer / ErrorlessIlluminaSequencer    src/sequencers.hpp:25-72
// with the Rcpp types replaced by standard ones (SEXP NULL -> std::nullopt / nullptr,
// Rcpp::List -> SeqResult).  Everything below the argument handling is libpcs_seq.so
// (include/pcs_seq.h); INTEGRATION.md shows the same calls made from the Rcpp layer.
// Header-only; link with -lpcs_seq.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <filesystem>
#include <functional>
#include <limits>
#include <map>
#include <memory>
#include <optional>
#include <ostream>
#include <random>
#include <set>
#include <stdexcept>
#include <string>
#include <variant>
#include <vector>

#include "../../include/pcs_seq.h"

namespace process_b200 {

// ------------------------------------------------------------------ sequencers
class ErrorlessIlluminaSequencer {
 public:
  double get_error_rate() const { return 0; }
  // show(): src/sequencers.cpp:25-31 (model and platform names are RACES'; "[UNVERIFIED-RACES]" wording)
  void show(std::ostream& os) const { os << "Errorless Illumina (platform: \"ILLUMINA\")" << std::endl; }
};

class BasicIlluminaSequencer {
  double error_rate;
  bool random_quality_scores;

 public:
  // build_sequencer(): src/sequencers.cpp:80-106
  BasicIlluminaSequencer(const double error_rate, const bool random_quality_scores = true)
      : error_rate(error_rate), random_quality_scores(random_quality_scores) {
    if (!(error_rate >= 0))
      throw std::domain_error("The parameter \"error_rate\" must be a positive real number.");
  }
  const double& get_error_rate() const { return error_rate; }
  void set_error_rate(const double& e) { error_rate = e; }
  const bool& producing_random_scores() const { return random_quality_scores; }
  void set_random_scores(const bool& r) { random_quality_scores = r; }
  // show(): src/sequencers.cpp:44-78
  void show(std::ostream& os) const {
    os << "Basic Illumina (platform: \"ILLUMINA\" error rate: " << std::to_string(error_rate)
       << (random_quality_scores ? " random quality scores" : " constant quality scores") << ")" << std::endl;
  }
};

// sequencer = NULL | ErrorlessIlluminaSequencer | BasicIlluminaSequencer (src/seq_simulation.cpp:386-428)
using Sequencer = std::variant<std::monostate, ErrorlessIlluminaSequencer, BasicIlluminaSequencer>;

// ------------------------------------------------------------------ forest
// One SID row of the output: what RACES::Mutations::SID + SIDData carry (src/seq_simulation.cpp:52-90)
struct SIDRow {
  std::string ref, alt;
  std::set<std::string> causes;   // empty -> NA
  std::set<std::string> classes;  // "driver" | "passenger" | "germinal" | "preneoplastic"
};

// const view of a sampled cell for the labelling function (src/sampled_cell.hpp:27-56)
struct SampledCell {
  uint32_t cell_id;
  std::string sample, epistate, mutant, species;
  double birth_time = 0;
};

// The event-labelled forest in the layout of pcs_forest_desc, plus the strings that stay on the host.
struct PhylogeneticForest {
  std::vector<std::string> chr_names, sample_names;
  std::vector<uint32_t> chr_len;
  std::vector<uint8_t> chr_n_alleles;
  std::vector<int32_t> node_parent;
  std::vector<uint32_t> leaf_node, leaf_sample;
  std::vector<uint64_t> node_event_off;
  std::vector<uint8_t> ev_kind, ev_nature;
  std::vector<uint16_t> ev_chr, ev_allele, ev_dest;
  std::vector<uint32_t> ev_pos, ev_len, ev_mut;
  std::vector<uint16_t> mut_chr;
  std::vector<uint32_t> mut_pos;
  std::vector<uint8_t> mut_ref_len, mut_alt_len;
  std::vector<uint32_t> germ_mut;
  std::vector<uint8_t> germ_allele_mask;
  std::vector<SIDRow> rows;                 // [n_mut]
  std::vector<SampledCell> cells;           // [n_leaves], optional (labelling)
  std::filesystem::path reference_path;

  const std::filesystem::path& get_reference_path() const { return reference_path; }

  pcs_forest_desc desc() const {
    pcs_forest_desc d{};
    d.n_chr = static_cast<uint32_t>(chr_len.size());
    d.chr_len = chr_len.data();
    d.chr_n_alleles = chr_n_alleles.data();
    d.n_nodes = static_cast<uint32_t>(node_parent.size());
    d.node_parent = node_parent.data();
    d.n_samples = static_cast<uint32_t>(sample_names.size());
    d.n_leaves = static_cast<uint32_t>(leaf_node.size());
    d.leaf_node = leaf_node.data();
    d.leaf_sample = leaf_sample.data();
    d.n_events = ev_kind.size();
    d.node_event_off = node_event_off.data();
    d.ev_kind = ev_kind.data();
    d.ev_chr = ev_chr.data();
    d.ev_pos = ev_pos.data();
    d.ev_len = ev_len.data();
    d.ev_allele = ev_allele.data();
    d.ev_dest = ev_dest.data();
    d.ev_mut = ev_mut.data();
    d.ev_nature = ev_nature.data();
    d.n_mut = static_cast<uint32_t>(mut_pos.size());
    d.mut_chr = mut_chr.data();
    d.mut_pos = mut_pos.data();
    d.mut_ref_len = mut_ref_len.data();
    d.mut_alt_len = mut_alt_len.data();
    d.n_germline = germ_mut.size();
    d.germ_mut = germ_mut.data();
    d.germ_allele_mask = germ_allele_mask.data();
    return d;
  }
};

// ------------------------------------------------------------------ result
struct SampleColumns {
  std::string name;                 // columns "<name>.occurrences", "<name>.coverage", "<name>.VAF"
  std::vector<int> occurrences, coverage;
  std::vector<double> VAF;
};

struct SeqParameters {  // the echo of src/seq_simulation.cpp:586-597 / :665-675
  std::optional<std::string> sequencer_name;
  double error_rate = 0;
  std::optional<bool> random_quality_scores;
  std::optional<std::string> reference_genome;
  std::optional<std::vector<std::string>> chromosomes;
  double coverage;
  int read_size, insert_size_mean, insert_size_stddev;
  std::string output_dir;
  bool write_SAM, update_SAM;
  double purity;
  bool with_normal_sample, preneoplastic;
  std::string filename_prefix, template_name_prefix;
  bool include_non_sequenced_mutations;
  int seed;
};

struct SeqResult {
  // "mutations" data frame
  std::vector<std::string> chr;
  std::vector<int> chr_pos;
  std::vector<std::string> ref, alt;
  std::vector<std::optional<std::string>> causes;
  std::vector<std::string> classes;
  std::vector<SampleColumns> samples;  // in sample-name order
  SeqParameters parameters;
  pcs_run_stats stats{};
};

namespace detail {

inline void pcs_check(int rc) {
  if (rc == PCS_OK) return;
  if (rc == PCS_ERR_INVALID) throw std::domain_error(pcs_last_error());
  throw std::runtime_error(pcs_last_error());
}

inline std::string join(const std::set<std::string>& S, const char sep = ';') {
  std::string out;
  for (const auto& s : S) {
    if (!out.empty()) out += sep;
    out += s;
  }
  return out;
}

inline std::string ordtostr(size_t i) {
  const char* suf = (i % 100 >= 10 && i % 100 <= 20) ? "th" : (i % 10 == 1 ? "st" : i % 10 == 2 ? "nd" : i % 10 == 3 ? "rd" : "th");
  return std::to_string(i) + suf;
}

// get_reference_genome(): src/seq_simulation.cpp:245-280
inline std::filesystem::path get_reference_genome(const PhylogeneticForest& forest,
                                                  const std::optional<std::string>& reference_genome) {
  if (!reference_genome) {
    const auto ref_genome = forest.get_reference_path();
    if (!std::filesystem::exists(ref_genome))
      throw std::runtime_error("The reference genome file \"" + ref_genome.string() +
                               "\" does not exists anymore. Please, re-build the mutation engine or use the "
                               "parameter \"reference_genome\".");
    return ref_genome;
  }
  if (!std::filesystem::exists(*reference_genome))
    throw std::runtime_error("The reference genome file \"" + *reference_genome + "\" does not exists.");
  return *reference_genome;
}

// get_random_seed<int>(): src/utility.hpp:41-64 (NULL -> a uniform int)
inline int get_random_seed(const std::optional<int>& seed) {
  if (seed) return *seed;
  std::random_device rd;
  return std::uniform_int_distribution<int>(std::numeric_limits<int>::min(), std::numeric_limits<int>::max())(rd);
}

// one context per process and device; forests stay resident between calls
struct Device {
  pcs_ctx* ctx = nullptr;
  std::map<const PhylogeneticForest*, pcs_forest*> forests;
  explicit Device(int id) { pcs_check(pcs_create(&ctx, id, nullptr)); }
  ~Device() {
    for (auto& kv : forests) pcs_forest_free(kv.second);
    if (ctx) pcs_destroy(ctx);
  }
  pcs_forest* resident(const PhylogeneticForest& f) {
    auto it = forests.find(&f);
    if (it != forests.end()) return it->second;
    pcs_forest_desc d = f.desc();
    pcs_forest* h = nullptr;
    pcs_check(pcs_forest_upload(ctx, &d, &h));
    forests.emplace(&f, h);
    return h;
  }
};

inline Device& device(int id = 0) {
  static std::map<int, std::unique_ptr<Device>> devs;
  auto& d = devs[id];
  if (!d) d = std::make_unique<Device>(id);
  return *d;
}

struct Call {
  const Sequencer* sequencer;
  std::optional<std::string> reference_genome;
  std::optional<std::vector<std::string>> chromosome_ids;
  double coverage;
  int read_size, insert_size_mean, insert_size_stddev;
  std::string output_dir;
  bool write_SAM, update_SAM_dir;
  double purity;
  bool with_normal_sample, preneoplastic, normal_only;
  std::string filename_prefix, template_name_prefix;
  bool include_non_sequenced_mutations;
  std::optional<int> seed;
};

inline SeqResult run(const PhylogeneticForest& forest, const Call& c, const std::vector<uint32_t>* groups,
                     const std::vector<std::string>& group_names) {
  const std::string ref_genome = get_reference_genome(forest, c.reference_genome).string();
  const int c_seed = get_random_seed(c.seed);

  pcs_seq_params P{};
  P.seed = c_seed;
  P.coverage = c.coverage;
  P.purity = c.purity;
  if (c.read_size < 1) throw std::domain_error("read_size must be in [1, 65535]");
  P.read_size = static_cast<uint32_t>(c.read_size);
  if (c.insert_size_mean < 0 || c.insert_size_stddev < 0) throw std::domain_error("The insert size must be non-negative.");
  P.insert_size_mean = static_cast<uint32_t>(c.insert_size_mean);
  P.insert_size_stddev = static_cast<uint32_t>(c.insert_size_stddev);
  P.with_normal_sample = c.with_normal_sample;
  P.preneoplastic_in_normal = c.preneoplastic;
  P.normal_only = c.normal_only;
  P.shard_count = 1;
  SeqParameters echo{};
  // sequencer dispatch: src/seq_simulation.cpp:375-429; echo: :453-515
  P.sequencer = PCS_SEQ_ERRORLESS;
  if (const auto* b = std::get_if<BasicIlluminaSequencer>(c.sequencer)) {
    P.sequencer = b->producing_random_scores() ? PCS_SEQ_BASIC_RANDOM : PCS_SEQ_BASIC_CONSTANT;
    P.error_rate = b->get_error_rate();
    echo.sequencer_name = "BasicIlluminaSequencer";
    echo.error_rate = b->get_error_rate();
    echo.random_quality_scores = b->producing_random_scores();
  } else if (std::holds_alternative<ErrorlessIlluminaSequencer>(*c.sequencer)) {
    echo.sequencer_name = "ErrorlessIlluminaSequencer";
  }
  // get_relevant_chr_set(): src/seq_simulation.cpp:299-352
  std::vector<uint8_t> mask;
  if (c.chromosome_ids) {
    mask.assign(forest.chr_names.size(), 0);
    for (const auto& name : *c.chromosome_ids) {
      auto it = std::find(forest.chr_names.begin(), forest.chr_names.end(), name);
      if (it == forest.chr_names.end()) throw std::domain_error("Unknown chromosome \"" + name + "\"");
      mask[it - forest.chr_names.begin()] = 1;
    }
    P.chr_mask = mask.data();
  }

  Device& dev = device(0);
  pcs_forest* fo = dev.resident(forest);
  if (!c.normal_only)
    pcs_check(pcs_forest_set_groups(fo, groups ? groups->data() : nullptr, static_cast<uint32_t>(group_names.size())));

  std::vector<std::string> names;
  if (c.normal_only) {
    names = {"normal_sample"};  // src/seq_simulation.cpp:650-652
  } else {
    names = group_names;
    if (c.with_normal_sample) names.push_back("normal_sample");  // :572-575
  }
  const size_t S = names.size(), M = forest.mut_pos.size();
  SeqResult res;
  // the rows of the data frame are assembled on the device (get_result_dataframe(): src/seq_simulation.cpp:52-181):
  // only the active rows' columns cross the link
  pcs_result* result = nullptr;
  struct ResultGuard {
    pcs_result*& r;
    ~ResultGuard() { pcs_result_free(r); }
  } result_guard{result};
  if (c.write_SAM) {
    // Mode::CREATE / Mode::UPDATE (src/seq_simulation.cpp:545-549); reads and tables come from one plan
    std::vector<const char*> chr_names, sample_names;
    for (const auto& n : forest.chr_names) chr_names.push_back(n.c_str());
    for (const auto& n : names) sample_names.push_back(n.c_str());
    pcs_check(pcs_forest_load_fasta(fo, ref_genome.c_str(), chr_names.data(), nullptr));
    std::vector<uint32_t> alt_off(M + 1, 0);
    std::string alt_bytes;
    for (size_t r = 0; r < M; ++r) {
      alt_bytes += forest.rows[r].alt;
      alt_off[r + 1] = static_cast<uint32_t>(alt_bytes.size());
    }
    pcs_check(pcs_forest_set_alt(fo, alt_off.data(), alt_bytes.c_str()));
    pcs_plan* plan = nullptr;
    pcs_check(pcs_plan_create(fo, &P, &plan));
    struct PlanGuard {
      pcs_plan* p;
      ~PlanGuard() { pcs_plan_free(p); }
    } guard{plan};
    std::vector<uint32_t> occ(S * M), cov(S * M);
    pcs_check(pcs_plan_run(plan, PCS_RUN_HOST_OUTPUT, occ.data(), cov.data(), &res.stats));
    pcs_sam_options opt{};
    opt.output_dir = c.output_dir.c_str();
    opt.filename_prefix = c.filename_prefix.c_str();
    opt.template_name_prefix = c.template_name_prefix.c_str();
    opt.chr_names = chr_names.data();
    opt.sample_names = sample_names.data();
    opt.update = c.update_SAM_dir ? 1 : 0;
    pcs_check(pcs_plan_write_sam(plan, &opt, nullptr));
    pcs_check(pcs_plan_result(plan, c.include_non_sequenced_mutations, &P, 1, &result));
  } else {
    pcs_check(pcs_simulate_result(fo, &P, c.include_non_sequenced_mutations, 1, &result, &res.stats));
  }

  uint32_t n = 0;
  pcs_check(pcs_result_info(result, &n, nullptr, nullptr));
  std::vector<uint32_t> rows(n);
  std::vector<size_t> order(S);
  for (size_t s = 0; s < S; ++s) order[s] = s;
  std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return names[a] < names[b]; });
  res.samples.resize(S);
  std::vector<int32_t*> occ_cols(S), cov_cols(S);
  std::vector<double*> vaf_cols(S);
  for (size_t k = 0; k < S; ++k) {  // columns in sample-name order, filled in place by the library
    SampleColumns& col = res.samples[k];
    col.name = names[order[k]];
    col.occurrences.resize(n);
    col.coverage.resize(n);
    col.VAF.resize(n);
    occ_cols[order[k]] = col.occurrences.data();
    cov_cols[order[k]] = col.coverage.data();
    vaf_cols[order[k]] = col.VAF.data();  // occurrences / coverage (:126); never covered: 0 (:129-131)
  }
  pcs_check(pcs_result_fetch(result, rows.data(), occ_cols.data(), cov_cols.data(), vaf_cols.data(), nullptr));
  for (uint32_t r : rows) {
    res.chr.push_back(forest.chr_names[forest.mut_chr[r]]);
    res.chr_pos.push_back(static_cast<int>(forest.mut_pos[r]));
    const SIDRow& sid = forest.rows[r];
    res.ref.push_back(sid.ref);
    res.alt.push_back(sid.alt);
    res.causes.push_back(sid.causes.empty() ? std::nullopt : std::optional<std::string>(join(sid.causes)));
    res.classes.push_back(join(sid.classes));
  }
  echo.reference_genome = c.reference_genome;
  echo.chromosomes = c.chromosome_ids;
  echo.coverage = c.coverage;
  echo.read_size = c.read_size;
  echo.insert_size_mean = c.insert_size_mean;
  echo.insert_size_stddev = c.insert_size_stddev;
  echo.output_dir = c.output_dir;
  echo.write_SAM = c.write_SAM;
  echo.update_SAM = c.update_SAM_dir;
  echo.purity = c.purity;
  echo.with_normal_sample = c.with_normal_sample;
  echo.preneoplastic = c.preneoplastic;
  echo.filename_prefix = c.filename_prefix;
  echo.template_name_prefix = c.template_name_prefix;
  echo.include_non_sequenced_mutations = c.include_non_sequenced_mutations;
  echo.seed = c_seed;
  res.parameters = echo;
  return res;
}

}  // namespace detail

using LabellingFunction = std::function<std::string(const SampledCell&)>;

// apply_FACS_labels()/split_by_labels(): src/seq_simulation.cpp:183-243
inline void apply_FACS_labels(const PhylogeneticForest& forest, const LabellingFunction* labelling,
                              std::vector<uint32_t>& groups, std::vector<std::string>& names) {
  names = forest.sample_names;
  groups.clear();
  if (!labelling) return;
  if (!*labelling) throw std::domain_error("The FACs_labelling_function must be a function.");
  names.clear();
  groups.assign(forest.leaf_node.size(), 0);
  std::map<std::pair<uint32_t, std::string>, uint32_t> index;
  for (uint32_t s = 0; s < forest.sample_names.size(); ++s)
    for (size_t l = 0; l < forest.leaf_node.size(); ++l) {
      if (forest.leaf_sample[l] != s) continue;
      SampledCell cell = l < forest.cells.size() ? forest.cells[l] : SampledCell{forest.leaf_node[l], "", "", "", ""};
      cell.cell_id = forest.leaf_node[l];
      cell.sample = forest.sample_names[s];
      const std::string label = (*labelling)(cell);
      auto key = std::make_pair(s, label);
      auto it = index.find(key);
      if (it == index.end()) {
        it = index.emplace(key, static_cast<uint32_t>(names.size())).first;
        names.push_back(label.empty() ? forest.sample_names[s] : forest.sample_names[s] + "_" + label);
      }
      groups[l] = it->second;
    }
}

inline SeqResult simulate_seq(const PhylogeneticForest& forest, const Sequencer& sequencer = {},
                              const std::optional<std::string>& reference_genome = std::nullopt,
                              const std::optional<std::vector<std::string>>& chromosome_ids = std::nullopt,
                              const double& coverage = 10, const int& read_size = 150, const int& insert_size_mean = 0,
                              const int& insert_size_stddev = 10, const std::string& output_dir = "ProCESS_SAM",
                              const bool& write_SAM = false, const bool& update_SAM_dir = false,
                              const LabellingFunction* FACS_labelling_function = nullptr, const double& purity = 1,
                              const bool& with_normal_sample = true, const bool& preneoplastic_in_normal = false,
                              const std::string& filename_prefix = "chr_", const std::string& template_name_prefix = "r",
                              const bool& include_non_sequenced_mutations = false,
                              const std::optional<int>& seed = std::nullopt) {
  if (!(purity >= 0 && purity <= 1)) throw std::domain_error("The purity must belong to the interval [0,1].");
  std::vector<uint32_t> groups;
  std::vector<std::string> names;
  apply_FACS_labels(forest, FACS_labelling_function, groups, names);
  detail::Call c{&sequencer, reference_genome, chromosome_ids, coverage, read_size, insert_size_mean,
                 insert_size_stddev, output_dir, write_SAM, update_SAM_dir, purity, with_normal_sample,
                 preneoplastic_in_normal, false, filename_prefix, template_name_prefix,
                 include_non_sequenced_mutations, seed};
  return detail::run(forest, c, FACS_labelling_function ? &groups : nullptr, names);
}

// defaults as the reference's: output_dir "ProCESS_normal_SAM", write_SAM = TRUE (src/sequencing.cpp:268-282)
inline SeqResult simulate_normal_seq(const PhylogeneticForest& forest, const Sequencer& sequencer = {},
                                     const std::optional<std::string>& reference_genome = std::nullopt,
                                     const std::optional<std::vector<std::string>>& chromosome_ids = std::nullopt,
                                     const double& coverage = 10, const int& read_size = 150,
                                     const int& insert_size_mean = 0, const int& insert_size_stddev = 10,
                                     const std::string& output_dir = "ProCESS_normal_SAM", const bool& write_SAM = true,
                                     const bool& update_SAM_dir = false, const bool& with_preneoplastic = false,
                                     const std::string& filename_prefix = "chr_",
                                     const std::string& template_name_prefix = "r",
                                     const bool& include_non_sequenced_mutations = false,
                                     const std::optional<int>& seed = std::nullopt) {
  detail::Call c{&sequencer, reference_genome, chromosome_ids, coverage, read_size, insert_size_mean,
                 insert_size_stddev, output_dir, write_SAM, update_SAM_dir, 1.0, false, with_preneoplastic, true,
                 filename_prefix, template_name_prefix, include_non_sequenced_mutations, seed};
  return detail::run(forest, c, nullptr, {});
}

}  // namespace process_b200
