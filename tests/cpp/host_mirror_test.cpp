// Test driver of the C++ host mirror (process_b200/csrc/process_seq.hpp) on the hand-computed
// micro forest of tests/golden/micro_forest.py.
//   host_mirror_test validate            argument handling / error behaviour, no GPU needed
//   host_mirror_test run <existing file> simulate on cuda:0, print the tables as TSV
#include <cstdio>
#include <cstring>
#include <iostream>
#include <sstream>

#include "../../process_b200/csrc/process_seq.hpp"

using namespace process_b200;

static PhylogeneticForest micro_forest() {
  PhylogeneticForest f;
  f.chr_names = {"1"};
  f.chr_len = {1000};
  f.chr_n_alleles = {2};
  f.node_parent = {-1, 0, 0};
  f.sample_names = {"s0", "s1"};
  f.leaf_node = {1, 2};
  f.leaf_sample = {0, 1};
  f.node_event_off = {0, 1, 2, 6};
  f.ev_kind = {0, 0, 1, 0, 0, 2};
  f.ev_chr = {0, 0, 0, 0, 0, 0};
  f.ev_pos = {0, 0, 700, 0, 0, 1};
  f.ev_len = {0, 0, 200, 0, 0, 150};
  f.ev_allele = {1, 0, 0, 2, 1, 0};
  f.ev_dest = {0, 0, 2, 0, 0, 0};
  f.ev_mut = {2, 4, 0, 6, 5, 0};
  f.ev_nature = {3, 1, 1, 1, 1, 1};
  f.mut_chr = {0, 0, 0, 0, 0, 0, 0, 0};
  f.mut_pos = {100, 200, 300, 302, 400, 600, 800, 905};
  f.mut_ref_len = {1, 1, 5, 1, 1, 1, 1, 1};
  f.mut_alt_len = {1, 1, 1, 1, 1, 4, 1, 1};
  f.germ_mut = {0, 1, 3, 7};
  f.germ_allele_mask = {1, 3, 2, 1};
  const char* cls[8] = {"germinal", "germinal", "preneoplastic", "germinal", "passenger", "passenger", "passenger", "germinal"};
  for (int i = 0; i < 8; ++i) {
    SIDRow r;
    r.ref = i == 2 ? "GTTTT" : "A";
    r.alt = i == 2 ? "G" : (i == 5 ? "ACCC" : "C");
    r.classes = {cls[i]};
    if (i == 4 || i == 6) r.causes = {"SBS1"};
    f.rows.push_back(r);
  }
  return f;
}

template <class E, class F>
static bool throws(F&& fn, const char* needle) {
  try {
    fn();
  } catch (const E& e) {
    if (std::strstr(e.what(), needle)) return true;
    std::fprintf(stderr, "wrong message: %s\n", e.what());
    return false;
  } catch (...) {
    std::fprintf(stderr, "wrong exception type\n");
    return false;
  }
  std::fprintf(stderr, "no exception (expected: %s)\n", needle);
  return false;
}

int main(int argc, char** argv) {
  PhylogeneticForest f = micro_forest();
  if (argc >= 2 && std::string(argv[1]) == "validate") {
    bool ok = true;
    ok &= throws<std::domain_error>([] { BasicIlluminaSequencer s(-1e-3); }, "must be a positive real number");
    ok &= throws<std::runtime_error>([&] { simulate_seq(f); }, "does not exists anymore");
    ok &= throws<std::runtime_error>([&] { simulate_seq(f, {}, std::string("/nonexistent/ref.fa")); }, "does not exists.");
    f.reference_path = argv[0];
    ok &= throws<std::domain_error>([&] { simulate_seq(f, {}, std::nullopt, std::nullopt, 10, 150, 0, 10, "ProCESS_SAM", false, false, nullptr, 1.5); }, "purity");
    BasicIlluminaSequencer b(4e-3);
    ok &= b.get_error_rate() == 4e-3 && b.producing_random_scores();
    b.set_random_scores(false);
    ok &= !b.producing_random_scores() && ErrorlessIlluminaSequencer().get_error_rate() == 0;
    {
      std::ostringstream os;
      b.show(os);
      ErrorlessIlluminaSequencer().show(os);
      ok &= os.str() == "Basic Illumina (platform: \"ILLUMINA\" error rate: 0.004000 constant quality scores)\n"
                        "Errorless Illumina (platform: \"ILLUMINA\")\n";
    }
    std::vector<uint32_t> groups;
    std::vector<std::string> names;
    LabellingFunction lab = [](const SampledCell& c) { return c.cell_id == 1 ? std::string("") : std::string("B"); };
    apply_FACS_labels(f, &lab, groups, names);
    ok &= names == std::vector<std::string>{"s0", "s1_B"} && groups == std::vector<uint32_t>{0, 1};
    std::puts(ok ? "ok" : "FAILED");
    return ok ? 0 : 1;
  }
  if (argc >= 3 && std::string(argv[1]) == "run") {
    f.reference_path = argv[2];
    Sequencer seq = BasicIlluminaSequencer(1e-2, false);
    SeqResult r = simulate_seq(f, seq, std::nullopt, std::vector<std::string>{"1"}, 400.0, 20, 0, 10, "ProCESS_SAM",
                               false, false, nullptr, 0.8, true, true, "chr_", "r", false, 7);
    std::printf("chr\tchr_pos\tref\talt\tcauses\tclasses");
    for (const auto& s : r.samples) std::printf("\t%s.occurrences\t%s.coverage\t%s.VAF", s.name.c_str(), s.name.c_str(), s.name.c_str());
    std::printf("\n");
    for (size_t i = 0; i < r.chr.size(); ++i) {
      std::printf("%s\t%d\t%s\t%s\t%s\t%s", r.chr[i].c_str(), r.chr_pos[i], r.ref[i].c_str(), r.alt[i].c_str(),
                  r.causes[i] ? r.causes[i]->c_str() : "NA", r.classes[i].c_str());
      for (const auto& s : r.samples) std::printf("\t%d\t%d\t%.6f", s.occurrences[i], s.coverage[i], s.VAF[i]);
      std::printf("\n");
    }
    std::printf("#seed\t%d\tsequencer\t%s\treads\t%llu\n", r.parameters.seed, r.parameters.sequencer_name->c_str(),
                static_cast<unsigned long long>(r.stats.n_reads));
    const std::string sam_dir = argc >= 4 ? argv[3] : "ProCESS_normal_SAM";
    SeqResult n = simulate_normal_seq(f, seq, std::nullopt, std::nullopt, 400.0, 20, 0, 10, sam_dir, argc >= 4,
                                      false, true, "chr_", "r", true, 7);
    std::printf("#normal\t%zu\t%s\t%zu\n", n.samples.size(), n.samples[0].name.c_str(), n.chr.size());
    return 0;
  }
  std::fprintf(stderr, "usage: %s validate | run <reference file>\n", argv[0]);
  return 2;
}
