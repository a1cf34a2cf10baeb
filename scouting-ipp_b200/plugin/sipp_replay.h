// Offline evaluation replay: the follower plan for every map an experiment saved, in one batched pass on the GPU.
//
// The reference's evaluation walks an experiment folder and starts one planner process per saved map
// (data_analysis/evaluation.py:39-62,139-144 -> planexp_base_planners/src/planexp_eval_planner.cpp:201-303): load the map, insert it
// into a fresh planner (update_states), plan(), append a row to the results table (common/utils.h:367-373) and write the path
// (:375-386). Here the maps of a run are loaded by the host, uploaded into one context each and planned TOGETHER by sipp_batch_plan;
// rows and path files keep the reference's format so data_analysis/evaluation.py reads them unchanged.
//
// Saved-map format: what MapServerPlanExp::saveMap writes (map_server_planexp.cpp:489-535, eigen_csv.h) — "<prefix>_cost.csv"
// (map_cost_, the label per cell) and "<prefix>_occupancy.csv" (map_state_: 0 occupied, 1 free, 2 unknown), rows of ", " separated
// numbers. A cell whose state is unknown was never observed and is not ingested (MapServerPlanExp::loadMap keeps it unknown too).
#ifndef SIPP_REPLAY_H_
#define SIPP_REPLAY_H_

#include <string>
#include <vector>

#include "../../include/sipp.h"

namespace sipp_replay {

struct SavedMap {
  int W = 0, H = 0;
  std::vector<int32_t> labels;     // [H][W] map_cost_
  std::vector<uint8_t> observed;   // [H][W] 1 where map_state_ != MSR_UNKNOWN
};

// parses "<prefix>_cost.csv" + "<prefix>_occupancy.csv"; false + *error on a missing file, ragged rows or mismatching shapes
bool loadSavedMap(const std::string& prefix, SavedMap* out, std::string* error);
// the writer of the same two files (tests, and a host program that wants to save the GPU map server's state)
bool saveSavedMap(const std::string& prefix, const SavedMap& map, std::string* error);

struct ReplayResult {
  std::string map_name;
  double cost = -1.0;          // -1: target not reached (store_results' "Success" column is cost != -1)
  int status = 0;              // SIPP_PLAN_*
  std::vector<double> path_xy; // start -> goal, metres
};

// Plans every saved map of `prefixes` on `device` with the planner / map parameters of `desc` (desc.W / desc.H must match the files),
// `batch_size` maps per batched pass. Appends one row per map to `results_csv` (empty: none) and writes "<name>_path.csv" into
// `path_dir` (empty: none). Returns false + *error when a map cannot be loaded or a CUDA call fails.
bool replaySavedMaps(const std::vector<std::string>& prefixes, const sipp_map_desc& desc, int device, int batch_size, const std::string& results_csv,
                     const std::string& path_dir, std::vector<ReplayResult>* results, std::string* error);

}  // namespace sipp_replay

extern "C" {
// C entry for hosts without C++ linkage (tests/test_gpu_replay.py drives it through ctypes): prefixes = n NUL-terminated strings;
// cost / status: n entries. Returns 0 on success; the message of a failure is copied into err (err_cap bytes).
int sipp_replay_saved_maps(const char* const* prefixes, int32_t n, const sipp_map_desc* desc, int32_t device, int32_t batch_size,
                           const char* results_csv, const char* path_dir, double* cost, int32_t* status, char* err, int32_t err_cap);
int sipp_replay_save_map(const char* prefix, int32_t W, int32_t H, const int32_t* labels, const uint8_t* observed, char* err, int32_t err_cap);
// loads one saved map; labels / observed take up to cap cells (may be NULL to query the shape only)
int sipp_replay_load_map(const char* prefix, int32_t* W, int32_t* H, int32_t* labels, uint8_t* observed, int64_t cap, char* err, int32_t err_cap);
}

#endif  // SIPP_REPLAY_H_
