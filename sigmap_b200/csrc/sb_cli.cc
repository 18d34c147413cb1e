// sb_cli.cc -- `sigmap` command-line driver on top of the C ABI: the drop-in for the
// reference's `sigmap -i` / `sigmap -m` (flags of sigmap.cc:1331-1377, same files in and
// out).  Host-side only: parses flags, reads FASTA / .pt / BLOW5, calls smb_map_reads, writes
// the modified PAF.  Extra flags: --gpu N (device ordinal).  `-t` is accepted for
// compatibility (the GPU path does not use host mapping threads).
//
// Differences kept on purpose and documented in DESIGN.md: `-i` writes <prefix>.pt only (the
// device index is rebuilt from the point cloud at load time; nanoflann's <prefix>.si is
// neither needed nor produced); FAST5 input is not supported (BLOW5 only).
#include <dirent.h>
#include <sys/stat.h>
#include <sys/time.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/sigmap_b200.h"

namespace {

double now() {
  struct timeval tp;
  gettimeofday(&tp, nullptr);
  return tp.tv_sec + tp.tv_usec * 1e-6;
}

[[noreturn]] void die(const std::string &msg) {  // ExitWithMessage, utils.h:67
  fprintf(stderr, "%s\n", msg.c_str());
  exit(-1);
}

bool is_dir(const std::string &p) {
  struct stat st;
  return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}

void collect_blow5(const std::string &dir, std::vector<std::string> &out, int depth = 0) {
  DIR *d = opendir(dir.c_str());
  if (!d) return;
  std::vector<std::string> names;
  while (struct dirent *e = readdir(d)) names.push_back(e->d_name);
  closedir(d);
  std::sort(names.begin(), names.end());
  for (const std::string &n : names) {
    if (n == "." || n == "..") continue;
    const std::string p = dir + "/" + n;
    if (is_dir(p)) {
      if (depth < 1) collect_blow5(p, out, depth + 1);  // signal_batch.cc:36-47: one level down
    } else if (p.find(".blow5") != std::string::npos) {
      out.push_back(p);
    } else if (p.find(".fast5") != std::string::npos) {
      fprintf(stderr, "warning: FAST5 input is not supported by this build, skipping %s\n", p.c_str());
    }
  }
}

struct Args {
  bool index = false, map = false, help = false;
  std::string ref, model, ref_index, sig_dir, output;
  int dimension = 6, max_leaf = 20, threads = 1, gpu = 0;
  smb_params prm;
};

const char *kHelp =
    "Map ONT raw signal data (B200 build)\nUsage:\n  sigmap [OPTION...]\n\n"
    " Indexing options:\n  -i, --build-index      Build spatial index for reference\n"
    "  -d, --dimension INT    Dimension of spatial index [6]\n"
    "  -l, --max-leaf INT     Max leaf of spatial index [20]\n\n"
    " Mapping options:\n  -m, --map              Map signal data\n"
    "      --step-size INT    Seeding step size in reads [2]\n"
    "  -t, --num-threads INT  # threads for mapping [1] (accepted, unused on GPU)\n"
    "      --gpu INT          CUDA device ordinal [0]\n\n"
    " Input options:\n  -r, --ref FILE         Reference file\n  -p, --pore-model FILE  Pore model file\n"
    "  -x, --ref-index FILE   Reference index file\n  -s, --sig-dir DIR      Signal data directory\n\n"
    " Output options:\n  -o, --output arg       Output file\n\n"
    " Development options:\n      --search-radius FLT            Search radius for each seed [0.08]\n"
    "      --max-num-chunks INT           Max # chunks before stop trying to map a read [30]\n"
    "      --min-num-anchors INT          Min # anchors to stop mapping [10]\n"
    "      --min-num-anchors-output INT   Min # anchors to output mappings [10]\n"
    "      --stop-mapping FLOAT           best/second-best chaining score to stop mapping [1.4]\n"
    "      --stop-mapping-output FLOAT    best/second-best chaining score to output mappings [1.2]\n"
    "      --stop-mapping-mean FLOAT      best/mean chaining score to stop mapping [5]\n"
    "      --stop-mapping-mean-output FLOAT best/mean chaining score to output mappings [5]\n"
    "  -h, --help                         Print help\n";

Args parse(int argc, char **argv) {
  Args a;
  smb_default_params(&a.prm);
  auto need = [&](int &i) -> const char * {
    if (i + 1 >= argc) die(std::string("Option ") + argv[i] + " is missing an argument");
    return argv[++i];
  };
  for (int i = 1; i < argc; ++i) {
    std::string o = argv[i], v;
    size_t eq = o.find('=');
    bool has_v = false;
    if (o.rfind("--", 0) == 0 && eq != std::string::npos) {
      v = o.substr(eq + 1);
      o = o.substr(0, eq);
      has_v = true;
    }
    auto val = [&]() -> std::string { return has_v ? v : std::string(need(i)); };
    if (o == "-i" || o == "--build-index") a.index = true;
    else if (o == "-m" || o == "--map") a.map = true;
    else if (o == "-h" || o == "--help") a.help = true;
    else if (o == "-d" || o == "--dimension") a.dimension = atoi(val().c_str());
    else if (o == "-l" || o == "--max-leaf") a.max_leaf = atoi(val().c_str());
    else if (o == "-t" || o == "--num-threads") a.threads = atoi(val().c_str());
    else if (o == "--gpu") a.gpu = atoi(val().c_str());
    else if (o == "-r" || o == "--ref") a.ref = val();
    else if (o == "-p" || o == "--pore-model") a.model = val();
    else if (o == "-x" || o == "--ref-index") a.ref_index = val();
    else if (o == "-s" || o == "--sig-dir") a.sig_dir = val();
    else if (o == "-o" || o == "--output") a.output = val();
    else if (o == "--step-size") a.prm.step_size = atoi(val().c_str());
    else if (o == "--search-radius") a.prm.search_radius = (float)atof(val().c_str());
    else if (o == "--max-num-chunks") a.prm.max_num_chunks = atoi(val().c_str());
    else if (o == "--min-num-anchors") a.prm.min_num_anchors = atoi(val().c_str());
    else if (o == "--min-num-anchors-output") a.prm.min_num_anchors_output = atoi(val().c_str());
    else if (o == "--stop-mapping") a.prm.stop_mapping = (float)atof(val().c_str());
    else if (o == "--stop-mapping-output") a.prm.stop_mapping_output = (float)atof(val().c_str());
    else if (o == "--stop-mapping-mean") a.prm.stop_mapping_mean = (float)atof(val().c_str());
    else if (o == "--stop-mapping-mean-output") a.prm.stop_mapping_mean_output = (float)atof(val().c_str());
    else die("Option '" + o + "' does not exist");
  }
  return a;
}

int build_index(const Args &a) {
  if (a.ref.empty()) die("No reference file specified!");
  if (a.model.empty()) die("No pore model file specified!");
  if (a.output.empty()) die("No output file specified!");
  if (a.dimension != SMB_DIM) die("Only dimension 6 is supported by this build");
  fprintf(stderr, "Dimension: %d, max leaf: %d\nReference file: %s\nPore model file: %s\nOutput file: %s\n",
          a.dimension, a.max_leaf, a.ref.c_str(), a.model.c_str(), a.output.c_str());
  double t0 = now();
  std::vector<float> mean(4096), stdv(4096);
  if (smbh_pore_model_load(a.model.c_str(), mean.data(), stdv.data())) die("Cannot load pore model!");
  smbh_fasta fa;
  if (smbh_fasta_load(a.ref.c_str(), &fa)) die("Cannot find sequence file!");
  size_t n = smbh_build_point_cloud(fa.seqs, fa.lengths, fa.n, mean.data(), nullptr, nullptr);
  std::vector<uint64_t> pos(n);
  std::vector<float> val(n);
  smbh_build_point_cloud(fa.seqs, fa.lengths, fa.n, mean.data(), pos.data(), val.data());
  fprintf(stderr, "Collected %zu points.\n", n);
  if (smbh_pt_write(a.output.c_str(), pos.data(), val.data(), n, a.dimension, a.max_leaf))
    die("Cannot write index file!");
  smbh_fasta_free(&fa);
  fprintf(stderr, "Built index successfully in %fs.\n", now() - t0);
  return 0;
}

int map_reads(const Args &a) {
  fprintf(stderr, "Number of threads: %d\n", a.threads);
  if (a.ref.empty()) die("No reference file specified!");
  if (a.model.empty()) die("No pore model file specified!");
  if (a.ref_index.empty()) die("No reference index file specified!");
  if (a.sig_dir.empty()) die("No signal data directory specified!");
  if (a.output.empty()) die("No output file specified!");
  fprintf(stderr, "Reference file: %s\nPore model file: %s\nReference index file: %s\nSignal directory: %s\nOutput file: %s\n",
          a.ref.c_str(), a.model.c_str(), a.ref_index.c_str(), a.sig_dir.c_str(), a.output.c_str());
  if (!is_dir(a.sig_dir)) die("Signal directory is in valid!");
  double t0 = now();
  std::vector<std::string> files;
  collect_blow5(a.sig_dir, files);
  smbh_reads reads;
  memset(&reads, 0, sizeof reads);
  for (const std::string &f : files)
    if (smbh_blow5_read(f.c_str(), &reads)) die("Error in opening file " + f);
  fprintf(stderr, "Loaded %zu reads in %fs.\n", reads.n, now() - t0);
  smbh_fasta fa;
  if (smbh_fasta_load(a.ref.c_str(), &fa)) die("Cannot find sequence file!");
  smb_ctx *ctx = nullptr;
  if (smb_create(&ctx, a.gpu)) die(std::string("smb_create: ") + smb_last_error(nullptr));
  t0 = now();
  if (smb_index_load(ctx, a.ref_index.c_str())) die(std::string("smb_index_load: ") + smb_last_error(ctx));
  smb_index_set_contigs(ctx, fa.lengths, fa.n);
  fprintf(stderr, "Loaded index successfully in %fs.\n", now() - t0);
  std::vector<smb_mapping> rows(reads.n ? reads.n : 1);
  t0 = now();
  uint64_t zero_off[1] = {0};
  if (smb_map_reads(ctx, reads.raw, reads.n ? reads.read_off : zero_off, reads.digitisation, reads.range,
                    reads.offset, reads.n, &a.prm, rows.data()))
    die(std::string("smb_map_reads: ") + smb_last_error(ctx));
  const double dt = now() - t0;
  fprintf(stderr, "Finished mapping in %f, # reads: %zu\n", dt, reads.n);
  smb_stats st;
  smb_stats_get(ctx, &st);
  fprintf(stderr, "GPU: %.3f ms kernels (events %.3f, search %.3f, sort %.3f, chain %.3f), %llu samples, %llu queries, %llu hits\n",
          st.ms_total, st.ms_events, st.ms_search, st.ms_sort, st.ms_chain, (unsigned long long)st.samples,
          (unsigned long long)st.queries, (unsigned long long)st.hits);
  // rows grouped by contig (unmapped under contig 0), arrival order inside: sigmap.cc:197-241
  FILE *out = fopen(a.output.c_str(), "w");
  if (!out) die("Cannot open output file!");
  const double mt = reads.n ? dt * 1000.0 / reads.n : 0.0;
  std::vector<char> line(4096);
  for (uint32_t c = 0; c < std::max(fa.n, 1u); ++c) {
    for (size_t r = 0; r < reads.n; ++r) {
      const smb_mapping &m = rows[r];
      const uint32_t bin = m.mapped ? m.contig : 0;
      if (bin != c) continue;
      const char *cname = m.mapped && m.contig < fa.n ? fa.names[m.contig] : "*";
      const uint32_t clen = m.mapped && m.contig < fa.n ? fa.lengths[m.contig] : 0;
      smbh_format_paf(&m, reads.names[r], cname, clen, mt, line.data(), line.size());
      fputs(line.data(), out);
    }
  }
  fclose(out);
  smb_destroy(ctx);
  smbh_reads_free(&reads);
  smbh_fasta_free(&fa);
  return 0;
}

}  // namespace

int main(int argc, char **argv) {
  Args a = parse(argc, argv);
  if (a.index) return build_index(a);
  if (a.map) return map_reads(a);
  fputs(kHelp, stderr);
  return 0;
}
