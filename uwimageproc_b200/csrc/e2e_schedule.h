// e2e_schedule.h - sub-batch schedule of the host-buffer chain (uwip_chain_bgr8, cabi.cu).  Plain C++: tests/test_host_logic.py
// compiles it with g++ and checks the schedules on the CPU.
//
// Only the first upload and the last download are not hidden behind compute, so the schedule ramps up from a few frames,
// runs a body of three-wave sub-batches and ramps down again:
//   4K on 148 SMs, 256 frames:  4, 8, 16, 24, 37, 55, 55, 32, 16, 6, 3
// Sizes are whole CTA waves of the guided-filter marches where that matters (a launch costs ceil(frames x strips / SMs) CTA
// durations): u = one wave of the narrow layout (16 frames), two and three waves of the wide one (37, 55), two of the narrow
// one (32).  The ramp grows by at most 2 x (the upload of the next sub-batch hides behind this one: at 4K the PCIe link
// needs 0.5 ms per frame and direction, the chain 0.8) and about halves on the way down (the download of the one before
// hides behind this one).  Measured at 256 x 4K (scratch/e2e_sched.py): 1,118 frames/s with 16, 32 x 7, 16; 1,170 with
// this one; four-wave bodies (74 frames) lose again - the copies no longer hide (1,129).
#pragma once
#include <algorithm>
#include <vector>

static inline std::vector<int> e2e_schedule(int n, int nb_max, int sms, int strips_wide, int strips_narrow) {
  std::vector<int> sizes;
  if (n <= 0) return sizes;
  nb_max = std::max(1, nb_max);
  auto wide = [&](int k) { return std::max(1, k * sms / std::max(1, strips_wide)); };
  auto narrow = [&](int k) { return std::max(1, k * sms / std::max(1, strips_narrow)); };
  const int u = narrow(1);
  int up[6] = {std::max(1, u / 4), std::max(1, u / 2), u, std::max(1, 3 * u / 2), wide(2), wide(3)};
  for (int& v : up) v = std::min(v, nb_max);
  auto down_from = [&](int peak) {
    std::vector<int> d;
    int t = peak * 3 / 5;                 // 55 -> 33
    if (t >= narrow(2)) t = narrow(2);    // -> 32: two waves of the narrow layout
    else if (t >= u) t = u;
    while (t >= 1) {
      d.push_back(t);
      if (t <= 3) break;
      t = (t > u) ? u : (t == u ? std::max(1, 3 * u / 8) : std::max(1, t / 2));
    }
    return d;
  };
  // the highest ramp whose way up and down fits into n frames
  int levels = 0;
  std::vector<int> down;
  for (int k = 6; k >= 1; k--) {
    if (k > 1 && up[k - 1] <= up[k - 2]) continue;   // capped by the workspace (or a tiny wave): not a level of its own
    std::vector<int> d = down_from(up[k - 1]);
    long sum = 0;
    for (int i = 0; i < k; i++) sum += up[i];
    for (int v : d) sum += v;
    if (sum <= n) { levels = k; down = d; break; }
  }
  if (levels == 0) {   // a handful of frames: two halves, so that one copy overlaps one compute
    if (n > 1) { sizes.push_back(std::min(nb_max, (n + 1) / 2)); }
    int rem = n - (sizes.empty() ? 0 : sizes[0]);
    while (rem > 0) { int m = std::min(rem, nb_max); sizes.push_back(m); rem -= m; }
    return sizes;
  }
  long used = 0;
  for (int i = 0; i < levels; i++) { sizes.push_back(up[i]); used += up[i]; }
  for (int v : down) used += v;
  const int peak = up[levels - 1];
  int rem = n - (int)used;
  while (rem >= peak) { sizes.push_back(peak); rem -= peak; }
  // the rest goes where the pipeline is full: a few frames join the last peak sub-batch, more become one of their own before it
  if (rem > 0 && 2 * rem < peak && sizes.back() + rem <= nb_max) sizes.back() += rem;
  else if (rem > 0) sizes.insert(sizes.begin() + (levels - 1), rem);
  for (int v : down) sizes.push_back(v);
  return sizes;
}
