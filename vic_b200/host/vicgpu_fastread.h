// vicgpu_fastread.h -- the reference's per-cell parameter lookups without their O(Ncell^2) scans.
//
// Three places of the reference's loader are quadratic in the number of cells and make its own main() unusable on the domains the
// device steps in seconds ("the reference cannot practically run 1 M cells"):
//   read_vegparam()          read_vegparam.c:113-128   rewind + scan from the top of the vegetation file for EVERY cell
//   read_snowband()          read_snowband.c:41-47     the same on the snow-band file (called from initializeCell, vicNl.c:333)
//   ProgramState::initGrid() get_global_param.c:79-104 smallest lat / lon distance over ALL PAIRS of cells
// Nothing of the reference's parsing is restated here.  read_vegparam.c and read_snowband.c are compiled from where they lie with
//   #define read_vegparam vicref_read_vegparam_at / read_snowband vicref_read_snowband_at / rewind vicgpu_reader_seek
// (force-included oracle/shim/vicgpu_seek_shim.h; oracle/Makefile, objects *_seek.o): their one rewind() becomes a call of the hook below, which positions the stream on the wanted
// cell's record, found in an index built by ONE pass over the file that skips records exactly as the reference's scan does (same
// fscanf / fgets calls, same buffer sizes).  The reference's loop then finds its cell in the first record it looks at.  A cell that
// is not in the index (or a malformed file) falls back to the plain rewind, i.e. to the reference's own scan and its own messages.
// Like the reference's scan, the index keeps the FIRST record of a cell number.
//
// vicgpu_init_grid() gives initGrid()'s results from the sorted distinct coordinates: the smallest non-zero |a - b| over all pairs
// is the smallest difference of two neighbours in sorted order (rounding is monotonic), computed in the coordinates' own type (float).
#ifndef VICGPU_FASTREAD_H
#define VICGPU_FASTREAD_H
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <map>
#include <sys/stat.h>
#include <unordered_map>
#include <vector>
#include "vicNl.h"

// the reference's functions under the names the -D compile gives them
int vicref_read_vegparam_at(FILE*, cell_info_struct&, const ProgramState*);
void vicref_read_snowband_at(FILE*, soil_con_struct*, const int);
extern "C" void vicgpu_reader_seek(FILE* f);  // stands where the reference calls rewind()

namespace vicgpu_fastread {

struct RecordIndex {
  std::unordered_map<int, long> first;  // cell number -> offset of its first record
  dev_t dev = 0;                        // the file the index was built from: a stream address can be reused by another file
  ino_t ino = 0;
  off_t size = -1;
  time_t mtime = 0;
  bool built_for(FILE* f) const {
    struct stat st;
    return fstat(fileno(f), &st) == 0 && st.st_dev == dev && st.st_ino == ino && st.st_size == size && st.st_mtime == mtime;
  }
  void stamp(FILE* f) {
    struct stat st;
    if (fstat(fileno(f), &st) == 0) { dev = st.st_dev; ino = st.st_ino; size = st.st_size; mtime = st.st_mtime; }
  }
};

inline long& seek_target() {
  static thread_local long t = -1;
  return t;
}

// one pass over the vegetation parameter file; records are skipped as read_vegparam.c:117-128 skips them
inline RecordIndex index_vegparam(FILE* f, int lines_per_tile) {
  RecordIndex ix;
  const long keep = ftell(f);
  rewind(f);
  char str[500];
  for (;;) {
    const long off = ftell(f);
    int vegcel, numHRUs;
    if (fscanf(f, "%d %d", &vegcel, &numHRUs) != 2 || numHRUs < 0) break;
    ix.first.emplace(vegcel, off);
    bool eof = false;
    for (int i = 0; i <= numHRUs * lines_per_tile; i++)
      if (fgets(str, 500, f) == NULL) { eof = true; break; }
    if (eof) break;
  }
  clearerr(f);
  fseek(f, keep < 0 ? 0 : keep, SEEK_SET);
  ix.stamp(f);
  return ix;
}

// one pass over the snow-band file; lines are skipped as read_snowband.c:44-47 skips them
inline RecordIndex index_snowband(FILE* f) {
  RecordIndex ix;
  const long keep = ftell(f);
  rewind(f);
  char line[MAXSTRING];
  for (;;) {
    const long off = ftell(f);
    int cell;
    if (fscanf(f, "%d", &cell) != 1) break;
    ix.first.emplace(cell, off);
    if (fgets(line, MAXSTRING, f) == NULL) break;
  }
  clearerr(f);
  fseek(f, keep < 0 ? 0 : keep, SEEK_SET);
  ix.stamp(f);
  return ix;
}

inline std::map<FILE*, RecordIndex>& indices() {
  static std::map<FILE*, RecordIndex> m;
  return m;
}

// read_vegparam() with the scan replaced by a lookup; same arguments, same results, same messages
inline int read_vegparam_indexed(FILE* vegparam, cell_info_struct& cell, const ProgramState* state) {
  auto it = indices().find(vegparam);
  if (it != indices().end() && !it->second.built_for(vegparam)) {  // another (or a changed) file behind the same stream address
    indices().erase(it);
    it = indices().end();
  }
  if (it == indices().end()) it = indices().emplace(vegparam, index_vegparam(vegparam, state->options.VEGPARAM_LAI ? 2 : 1)).first;
  auto rec = it->second.first.find(cell.soil_con.gridcel);
  seek_target() = rec == it->second.first.end() ? -1 : rec->second;
  const int n = vicref_read_vegparam_at(vegparam, cell, state);
  seek_target() = -1;
  return n;
}

inline void read_snowband_indexed(FILE* snowband, soil_con_struct* soil_con, const int num_elevation_snow_bands) {
  if (num_elevation_snow_bands > 1 && snowband) {
    auto it = indices().find(snowband);
    if (it != indices().end() && !it->second.built_for(snowband)) {
      indices().erase(it);
      it = indices().end();
    }
    if (it == indices().end()) it = indices().emplace(snowband, index_snowband(snowband)).first;
    auto rec = it->second.first.find(soil_con->gridcel);
    seek_target() = rec == it->second.first.end() ? -1 : rec->second;
  }
  vicref_read_snowband_at(snowband, soil_con, num_elevation_snow_bands);
  seek_target() = -1;
}

template <class T>
static double smallest_step(std::vector<T>& v) {
  double step = INT_MAX;  // get_global_param.c:74-75
  std::sort(v.begin(), v.end());
  for (size_t i = 1; i < v.size(); i++) {
    const double d = std::abs(v[i] - v[i - 1]);  // in T, as the reference's std::abs(lat_i - lat_n)
    if (d != 0 && d < step) step = d;
  }
  return step;
}

// ProgramState::initGrid(), get_global_param.c:61-109, in O(N log N)
inline void init_grid(global_param_struct& gp, const std::vector<cell_info_struct>& cells) {
  if (cells.size() == 0) throw VICException("Error: cannot run the model with no cells! Make sure that some cells are enabled.");
  if (cells.size() == 1) {
    gp.gridNumLatDivisions = 1;
    gp.gridNumLonDivisions = 1;
    gp.gridStartLat = cells[0].soil_con.lat;
    gp.gridStartLon = cells[0].soil_con.lng;
    gp.gridStepLat = 0;
    gp.gridStepLon = 0;
    return;
  }
  std::vector<decltype(cells[0].soil_con.lat)> lat(cells.size());
  std::vector<decltype(cells[0].soil_con.lng)> lng(cells.size());
  for (size_t i = 0; i < cells.size(); i++) {
    lat[i] = cells[i].soil_con.lat;
    lng[i] = cells[i].soil_con.lng;
  }
  gp.gridStepLat = smallest_step(lat);  // sorts
  gp.gridStepLon = smallest_step(lng);
  gp.gridStartLat = lat.front();
  gp.gridEndLat = lat.back();
  gp.gridStartLon = lng.front();
  gp.gridEndLon = lng.back();
  gp.gridNumLatDivisions = ((gp.gridEndLat - gp.gridStartLat) / gp.gridStepLat) + 1;
  gp.gridNumLonDivisions = ((gp.gridEndLon - gp.gridStartLon) / gp.gridStepLon) + 1;
}

}  // namespace vicgpu_fastread
#endif
