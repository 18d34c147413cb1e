// oracle/ref_harness.cc -- TEST INFRASTRUCTURE ONLY.
//
// A C-ABI shim that only CALLS the unmodified reference's public entry points
// so that stage outputs (pA samples, per-chunk features, radius-search hit
// sets, chains) can be dumped from Python and compared with oracle/sigmap_oracle.c
// and with the CUDA path.  It is linked against the reference objects compiled
// by oracle/Makefile (strict-FP build) into oracle/_ref/libsigmap_ref_stage.so.
//
// "sigmap.h" is included FIRST on purpose: the float/double overloads of
// abs/sqrt/fabs that event.h and sigmap.cc see depend on include order
// (SURVEY.md Q5/Q6); everything arithmetic stays inside the reference objects,
// this file never re-implements any of it.
#include "sigmap.h"
#include "spatial_index.h"

#include <slow5/slow5.h>

#include <cstring>
#include <string>
#include <vector>

namespace {
struct OpenIndex : public sigmap::SpatialIndex {
  OpenIndex(const std::string &prefix)
      : sigmap::SpatialIndex(1000, std::vector<int>(1000, 5000), prefix) {}
  // reference: protected members spatial_index.h:104-115
  SigmapAdaptor<float> *tree() { return spatial_index_; }
  std::vector<Point> &cloud() { return point_cloud_; }
  int dim() const { return dimension_; }
};
struct ChainState {
  std::vector<sigmap::SignalAnchorChain> chains;
};
}  // namespace

extern "C" {

// signal_batch.cc:182-210 (SignalBatch::AddSignal(slow5_rec_t*)): raw -> pA,
// range filter, compaction.  Returns the number of kept samples.
size_t ref_raw_to_pa(const int16_t *raw, size_t n, double digitisation,
                     double offset, double range, float *out) {
  slow5_rec_t rec;
  memset(&rec, 0, sizeof(rec));
  char id[] = "harness";
  rec.read_id = id;
  rec.read_id_len = (uint16_t)strlen(id);
  rec.digitisation = digitisation;
  rec.offset = offset;
  rec.range = range;
  rec.sampling_rate = 4000;
  rec.len_raw_signal = n;
  rec.raw_signal = const_cast<int16_t *>(raw);
  sigmap::SignalBatch batch;
  batch.AddSignal(&rec);
  const sigmap::Signal &s = batch.GetSignalAt(0);
  memcpy(out, s.signal_values.data(), s.signal_values.size() * sizeof(float));
  return s.signal_values.size();
}

// sigmap.cc:1048 Sigmap::GenerateEvents on pa[0..n).  Returns #features.
size_t ref_generate_events(const float *pa, size_t n, float *features,
                           float *stdvs, size_t cap) {
  sigmap::Signal sig;
  sig.signal_values.assign(pa, pa + n);
  sigmap::Sigmap sm;
  std::vector<float> f, s;
  sm.GenerateEvents(0, n, sig, f, s);
  size_t m = f.size() < cap ? f.size() : cap;
  if (features) memcpy(features, f.data(), m * sizeof(float));
  if (stdvs) memcpy(stdvs, s.data(), (s.size() < cap ? s.size() : cap) * sizeof(float));
  return f.size();
}

// event.h:226 DetectEvents on pa[0..n): raw events (mean, start, length) plus
// both t-statistic tracks and the peak list, for fine-grained parity tests.
size_t ref_detect_events(const float *pa, size_t n, float *tstat1, float *tstat2,
                         uint64_t *peaks, size_t *n_peaks, float *means,
                         uint64_t *starts, uint64_t *lengths, size_t cap) {
  std::vector<float> ps, pss, t1, t2;
  std::vector<size_t> pk;
  std::vector<sigmap::Event> ev;
  sigmap::DetectEvents(pa, n, sigmap::event_detection_defaults, ps, pss, t1, t2,
                       pk, ev);
  if (tstat1) memcpy(tstat1, t1.data(), t1.size() * sizeof(float));
  if (tstat2) memcpy(tstat2, t2.data(), t2.size() * sizeof(float));
  if (n_peaks) *n_peaks = pk.size();
  for (size_t i = 0; i < pk.size() && i < cap; ++i)
    if (peaks) peaks[i] = pk[i];
  for (size_t i = 0; i < ev.size() && i < cap; ++i) {
    if (means) means[i] = ev[i].mean;
    if (starts) starts[i] = ev[i].start;
    if (lengths) lengths[i] = ev[i].length;
  }
  return ev.size();
}

// spatial_index.cc:132 SpatialIndex::Load (needs <prefix>.pt and <prefix>.si)
void *ref_index_load(const char *prefix) {
  OpenIndex *idx = new OpenIndex(prefix);
  idx->Load();
  return idx;
}
void ref_index_free(void *h) { delete static_cast<OpenIndex *>(h); }
size_t ref_index_num_points(void *h) {
  return static_cast<OpenIndex *>(h)->cloud().size();
}
void ref_index_points(void *h, uint64_t *pos, float *val) {
  std::vector<Point> &c = static_cast<OpenIndex *>(h)->cloud();
  for (size_t i = 0; i < c.size(); ++i) {
    pos[i] = c[i].position;
    val[i] = c[i].value;
  }
}

// spatial_index.cc:366 index->radiusSearch(query, radius, out, sorted=false).
// Returns the number of hits (all of them; at most cap are copied, in the
// KD-tree traversal order the reference sees).
size_t ref_radius_search(void *h, const float *q, float radius, uint64_t *idx_out,
                         float *d2_out, size_t cap) {
  OpenIndex *idx = static_cast<OpenIndex *>(h);
  nanoflann::SearchParams params;
  params.sorted = false;
  std::vector<std::pair<size_t, float> > res;
  size_t n = idx->tree()->index->radiusSearch(q, radius, res, params);
  for (size_t i = 0; i < n && i < cap; ++i) {
    idx_out[i] = res[i].first;
    d2_out[i] = res[i].second;
  }
  return n;
}

void *ref_chain_state_new() { return new ChainState(); }
void ref_chain_state_free(void *s) { delete static_cast<ChainState *>(s); }
void ref_chain_state_clear(void *s) { static_cast<ChainState *>(s)->chains.clear(); }

// spatial_index.cc:276 SpatialIndex::GenerateChains; `state` carries the
// previous chunk's chains in and the new chains out, as StreamingMap does.
size_t ref_generate_chains(void *h, void *state, const float *features, size_t n,
                           uint32_t query_offset, int step, float radius,
                           size_t n_targets) {
  OpenIndex *idx = static_cast<OpenIndex *>(h);
  ChainState *st = static_cast<ChainState *>(state);
  std::vector<float> q(features, features + n), sd(n, 0.0f);
  idx->GenerateChains(q, sd, query_offset, step, radius, n_targets, st->chains);
  return st->chains.size();
}
size_t ref_chain_count(void *state) {
  return static_cast<ChainState *>(state)->chains.size();
}
// out7 = {contig, start, end, n_anchors, mapq, direction(1=+,0=-), anchors.size()}
void ref_chain_get(void *state, size_t i, float *score, uint32_t *out7) {
  const sigmap::SignalAnchorChain &c = static_cast<ChainState *>(state)->chains[i];
  *score = c.score;
  out7[0] = c.reference_sequence_index;
  out7[1] = c.start_position;
  out7[2] = c.end_position;
  out7[3] = c.num_anchors;
  out7[4] = c.mapq;
  out7[5] = c.direction == sigmap::Positive ? 1 : 0;
  out7[6] = (uint32_t)c.anchors.size();
}
void ref_chain_anchors(void *state, size_t i, uint32_t *target, uint32_t *query,
                       float *dist) {
  const sigmap::SignalAnchorChain &c = static_cast<ChainState *>(state)->chains[i];
  for (size_t a = 0; a < c.anchors.size(); ++a) {
    target[a] = c.anchors[a].target_position;
    query[a] = c.anchors[a].query_position;
    dist[a] = c.anchors[a].distance;
  }
}

}  // extern "C"
