// .history stream for voxcraft-viz: the text the reference writes with device printf
// (src/VX3/VX3_SimulationManager.cu:40-50 header, :70-114 frames).  Formatting happens on the host from a
// read-back of the simulation's state; the bytes go to the caller's vx3_history_cb.
#pragma once
#include <string>
#include <vector>

struct vx3_batch;

namespace vx3 {
struct HistoryWriter {
    // {{{setting}}} lines printed once per simulation when RecordStepSize > 0
    std::string header(const std::vector<int> &matid, const std::vector<float> &rgba, double vox_size) const;
};
} // namespace vx3

// one frame: "<<<Step%d Time:%f>>>...<<<>>>|[[[%d]]]...[[[]]]\n" for loop index j of simulation `sim`
static int history_frame(vx3_batch *b, int sim, long long j, double t, const vx3::HistoryWriter &hw, std::string &out);
