// see sipp_replay.h
#include "sipp_replay.h"

#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>

namespace sipp_replay {

namespace {

std::string num(double v);
std::string num(long long v);

// one CSV matrix of the reference's writer (Eigen IOFormat(StreamPrecision, DontAlignCols, ", ", "\n")); load_csv (eigen_csv.h:25-40)
// splits at ',' and parses every cell with std::stod, which skips the blank after the comma
bool loadMatrix(const std::string& path, std::vector<double>* values, int* rows, int* cols, std::string* error) {
  std::ifstream in(path);
  if (!in) { *error = "cannot open " + path; return false; }
  values->clear();
  *rows = 0;
  *cols = 0;
  std::string line;
  while (std::getline(in, line)) {
    if (line.empty()) continue;
    int c = 0;
    std::stringstream ls(line);
    std::string cell;
    while (std::getline(ls, cell, ',')) {
      try {
        values->push_back(std::stod(cell));
      } catch (...) {
        *error = path + ": cell '" + cell + "' in row " + std::to_string(*rows) + " is not a number";
        return false;
      }
      ++c;
    }
    if (*rows == 0) *cols = c;
    else if (c != *cols) { *error = path + ": row " + std::to_string(*rows) + " has " + std::to_string(c) + " cells, row 0 has " + std::to_string(*cols); return false; }
    ++*rows;
  }
  if (*rows == 0 || *cols == 0) { *error = path + " is empty"; return false; }
  return true;
}

template <class T>
bool saveMatrix(const std::string& path, const std::vector<T>& v, int rows, int cols, std::string* error) {
  std::ofstream out(path);
  if (!out) { *error = "cannot write " + path; return false; }
  for (int i = 0; i < rows; ++i) {
    for (int j = 0; j < cols; ++j) {
      if (j) out << ", ";
      out << num((long long)v[(size_t)i * cols + j]);
    }
    if (i + 1 < rows) out << "\n";      // Eigen's format puts the row separator BETWEEN rows
  }
  return (bool)out;
}

// numbers are formatted with snprintf, not with the stream's num_put facet: inside a host process that carries a second C++ runtime
// (a Python interpreter with bundled libraries) the facet lookup of this library's streams can fail silently. "%g" with 6 digits is what
// operator<<(double) prints by default, i.e. what the reference's store_results / store_path write.
std::string num(double v) {
  char buf[64];
  snprintf(buf, sizeof(buf), "%g", v);
  return buf;
}
std::string num(long long v) { return std::to_string(v); }

std::string baseName(const std::string& p) {
  const size_t k = p.find_last_of('/');
  return k == std::string::npos ? p : p.substr(k + 1);
}

struct CtxDeleter { void operator()(sipp_ctx* c) const { sipp_destroy(c); } };
struct BatchDeleter { void operator()(sipp_batch* b) const { sipp_batch_destroy(b); } };

}  // namespace

bool loadSavedMap(const std::string& prefix, SavedMap* out, std::string* error) {
  std::vector<double> cost, occ;
  int cr = 0, cc = 0, orows = 0, ocols = 0;
  if (!loadMatrix(prefix + "_cost.csv", &cost, &cr, &cc, error)) return false;
  if (!loadMatrix(prefix + "_occupancy.csv", &occ, &orows, &ocols, error)) return false;
  if (cr != orows || cc != ocols) {
    *error = prefix + ": cost map is " + std::to_string(cr) + " x " + std::to_string(cc) + ", occupancy map " + std::to_string(orows) + " x " + std::to_string(ocols);
    return false;
  }
  out->W = cc;
  out->H = cr;
  out->labels.resize(cost.size());
  out->observed.resize(cost.size());
  for (size_t i = 0; i < cost.size(); ++i) {
    out->observed[i] = occ[i] != (double)SIPP_STATE_UNKNOWN ? 1 : 0;     // MSR_UNKNOWN 2 (map_server_planexp.h:37-39)
    out->labels[i] = out->observed[i] ? (int32_t)cost[i] : 0;
  }
  return true;
}

bool saveSavedMap(const std::string& prefix, const SavedMap& map, std::string* error) {
  if (map.W < 1 || map.H < 1 || map.labels.size() != (size_t)map.W * map.H || map.observed.size() != map.labels.size()) {
    *error = "saveSavedMap: inconsistent map";
    return false;
  }
  std::vector<int> occ(map.labels.size());
  for (size_t i = 0; i < occ.size(); ++i) occ[i] = map.observed[i] ? SIPP_STATE_FREE : SIPP_STATE_UNKNOWN;
  return saveMatrix(prefix + "_cost.csv", map.labels, map.H, map.W, error) && saveMatrix(prefix + "_occupancy.csv", occ, map.H, map.W, error);
}

bool replaySavedMaps(const std::vector<std::string>& prefixes, const sipp_map_desc& desc, int device, int batch_size, const std::string& results_csv,
                     const std::string& path_dir, std::vector<ReplayResult>* results, std::string* error) {
  if (batch_size < 1) batch_size = 1;
  results->clear();
  results->resize(prefixes.size());
  const size_t nctx = prefixes.size() < (size_t)batch_size ? prefixes.size() : (size_t)batch_size;
  // the contexts are reused from pass to pass (sipp_map_clear = MapServerPlanExp::clearMap + a fresh planner)
  std::vector<std::unique_ptr<sipp_ctx, CtxDeleter>> ctxs;
  for (size_t k = 0; k < nctx; ++k) {
    sipp_ctx* c = nullptr;
    if (sipp_create(&desc, device, &c) != SIPP_OK) { *error = std::string("sipp_create: ") + sipp_last_error(nullptr); return false; }
    ctxs.emplace_back(c);
    sipp_set_incremental(c, 0);       // every saved map is planned from scratch, like a fresh planner process
  }
  std::vector<double> cost(nctx);
  std::vector<int32_t> status(nctx), len(nctx);
  for (size_t first = 0; first < prefixes.size(); first += nctx) {
    const size_t m = prefixes.size() - first < nctx ? prefixes.size() - first : nctx;
    std::vector<sipp_ctx*> raw;
    for (size_t k = 0; k < m; ++k) {
      SavedMap map;
      if (!loadSavedMap(prefixes[first + k], &map, error)) return false;
      if (map.W != desc.W || map.H != desc.H) {
        *error = prefixes[first + k] + ": map is " + std::to_string(map.W) + " x " + std::to_string(map.H) + ", the planner was set up for " +
                 std::to_string(desc.W) + " x " + std::to_string(desc.H);
        return false;
      }
      sipp_ctx* c = ctxs[k].get();
      if (sipp_map_clear(c) != SIPP_OK || sipp_map_upload_full(c, map.labels.data(), map.observed.data()) != SIPP_OK) {
        *error = prefixes[first + k] + ": " + sipp_last_error(c);
        return false;
      }
      raw.push_back(c);
    }
    sipp_batch* b = nullptr;
    if (sipp_batch_create(raw.data(), (int32_t)m, &b) != SIPP_OK) { *error = std::string("sipp_batch_create: ") + sipp_last_error(raw[0]); return false; }
    std::unique_ptr<sipp_batch, BatchDeleter> batch(b);
    if (sipp_batch_plan(b, cost.data(), status.data(), len.data()) != SIPP_OK) { *error = std::string("sipp_batch_plan: ") + sipp_batch_last_error(b); return false; }
    for (size_t k = 0; k < m; ++k) {
      ReplayResult& r = (*results)[first + k];
      r.map_name = baseName(prefixes[first + k]);
      r.cost = cost[k];
      r.status = status[k];
      if (len[k] > 0) {
        r.path_xy.resize((size_t)len[k] * 2);
        int32_t got = 0;
        if (sipp_path_download(raw[k], r.path_xy.data(), len[k], &got) != SIPP_OK) { *error = sipp_last_error(raw[k]); return false; }
        r.path_xy.resize((size_t)got * 2);
      }
    }
  }
  if (!results_csv.empty()) {      // store_results (common/utils.h:367-373): appended, one row per map
    std::ofstream out(results_csv, std::ios::app);
    if (!out) { *error = "cannot write " + results_csv; return false; }
    for (const ReplayResult& r : *results)
      out << r.map_name << "," << num(desc.width) << "," << num(desc.height) << "," << num(desc.start_x) << "," << num(desc.start_y) << "," << num(desc.goal_x) << ","
          << num(desc.goal_y) << "," << (r.cost != -1.0 ? "1" : "0") << "," << num(r.cost) << std::endl;
  }
  if (!path_dir.empty()) {         // store_path (:375-386)
    for (const ReplayResult& r : *results) {
      if (r.cost == -1.0) continue;
      std::ofstream out(path_dir + "/" + r.map_name + "_path.csv", std::ios::out | std::ios::trunc);
      if (!out) { *error = "cannot write into " + path_dir; return false; }
      out << "PositionX,PositionY,LocationX,LocationY" << std::endl << "m,m,px,px" << std::endl;
      for (size_t i = 0; i + 1 < r.path_xy.size(); i += 2) {
        // coordinate_to_pixle (common/utils.h): metres -> pixel indices
        const int px = (int)(r.path_xy[i] * desc.W / desc.width), py = (int)(r.path_xy[i + 1] * desc.H / desc.height);
        out << num(r.path_xy[i]) << "," << num(r.path_xy[i + 1]) << "," << num((long long)px) << "," << num((long long)py) << std::endl;
      }
    }
  }
  return true;
}

}  // namespace sipp_replay

extern "C" {

static void copy_err(const std::string& e, char* err, int32_t cap) {
  if (err && cap > 0) snprintf(err, (size_t)cap, "%s", e.c_str());
}

int sipp_replay_saved_maps(const char* const* prefixes, int32_t n, const sipp_map_desc* desc, int32_t device, int32_t batch_size, const char* results_csv,
                           const char* path_dir, double* cost, int32_t* status, char* err, int32_t err_cap) {
  if (!prefixes || n < 0 || !desc) { copy_err("bad arguments", err, err_cap); return SIPP_ERR_INVALID_ARG; }
  std::vector<std::string> p;
  for (int i = 0; i < n; ++i) p.emplace_back(prefixes[i] ? prefixes[i] : "");
  std::vector<sipp_replay::ReplayResult> res;
  std::string e;
  if (!sipp_replay::replaySavedMaps(p, *desc, device, batch_size, results_csv ? results_csv : "", path_dir ? path_dir : "", &res, &e)) {
    copy_err(e, err, err_cap);
    return SIPP_ERR_INVALID_ARG;
  }
  for (int i = 0; i < n; ++i) {
    if (cost) cost[i] = res[i].cost;
    if (status) status[i] = res[i].status;
  }
  return SIPP_OK;
}

int sipp_replay_save_map(const char* prefix, int32_t W, int32_t H, const int32_t* labels, const uint8_t* observed, char* err, int32_t err_cap) {
  if (!prefix || W < 1 || H < 1 || !labels || !observed) { copy_err("bad arguments", err, err_cap); return SIPP_ERR_INVALID_ARG; }
  sipp_replay::SavedMap m;
  m.W = W; m.H = H;
  m.labels.assign(labels, labels + (size_t)W * H);
  m.observed.assign(observed, observed + (size_t)W * H);
  std::string e;
  if (!sipp_replay::saveSavedMap(prefix, m, &e)) { copy_err(e, err, err_cap); return SIPP_ERR_INVALID_ARG; }
  return SIPP_OK;
}

int sipp_replay_load_map(const char* prefix, int32_t* W, int32_t* H, int32_t* labels, uint8_t* observed, int64_t cap, char* err, int32_t err_cap) {
  if (!prefix || !W || !H) { copy_err("bad arguments", err, err_cap); return SIPP_ERR_INVALID_ARG; }
  sipp_replay::SavedMap m;
  std::string e;
  if (!sipp_replay::loadSavedMap(prefix, &m, &e)) { copy_err(e, err, err_cap); return SIPP_ERR_INVALID_ARG; }
  *W = m.W;
  *H = m.H;
  const int64_t n = (int64_t)m.W * m.H;
  if ((labels || observed) && cap < n) { copy_err("buffer too small", err, err_cap); return SIPP_ERR_INVALID_ARG; }
  if (labels) memcpy(labels, m.labels.data(), (size_t)n * sizeof(int32_t));
  if (observed) memcpy(observed, m.observed.data(), (size_t)n);
  return SIPP_OK;
}

}  // extern "C"
