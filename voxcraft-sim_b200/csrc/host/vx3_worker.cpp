// Batch manager / node worker (include/vx3_worker.h).  Host code only; all device work goes through the C ABI
// of the engine (include/vx3_abi.h), so this file is also a reference consumer of that ABI.
#include <cuda_runtime_api.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/vx3_model.h"
#include "../../../include/vx3_worker.h"
#include "vx3_xml.h"

using namespace vx3;

void vx3_model_set_error(const std::string &msg);

namespace {
std::mutex g_out_mutex;

void history_to_stdout(void *, int, const char *bytes, size_t n) {
    std::lock_guard<std::mutex> lk(g_out_mutex);
    fwrite(bytes, 1, n, stdout);
}

std::string fmt_double(double v) { // boost::property_tree put<double>: stream with max_digits10
    char buf[64];
    if (std::isnan(v)) return std::signbit(v) ? "-nan" : "nan";
    snprintf(buf, sizeof(buf), "%.17g", v);
    return buf;
}
std::string join_path(const std::string &a, const std::string &b) {
    if (a.empty() || b.empty() || b[0] == '/') return b.empty() ? a : b;
    return a.back() == '/' ? a + b : a + "/" + b;
}
std::string base_name(const std::string &p) {
    std::string s = p;
    while (s.size() > 1 && s.back() == '/') s.pop_back();
    size_t k = s.find_last_of('/');
    return k == std::string::npos ? s : s.substr(k + 1);
}
} // namespace

extern "C" int vx3_write_report(const char *vxr_path, const char *input_dir, const vx3_result *r, int n) {
    return vx3_write_report_positions(vxr_path, input_dir, r, nullptr, n);
}

extern "C" int vx3_write_report_positions(const char *vxr_path, const char *input_dir, const vx3_result *r, const vx3_voxel_positions *vp, int n) {
    if (!vxr_path || !r || n <= 0) return VX3_ERR_INVALID;
    FILE *f = fopen(vxr_path, "w");
    if (!f) return VX3_ERR_INVALID;
    std::string s = "<?xml version=\"1.0\" encoding=\"utf-8\"?>\n<report>";
    s += "<inputdir>" + xml_escape(base_name(input_dir ? input_dir : "")) + "</inputdir>";
    s += "<bestfit><filename>" + xml_escape(r[0].name) + "</filename><fitness_score>" + fmt_double(r[0].fitness_score) + "</fitness_score></bestfit>";
    s += "<detail>";
    for (int i = 0; i < n; i++) {
        std::string nm = r[i].name;
        size_t dot = nm.find('.');
        if (dot != std::string::npos) nm = nm.substr(0, dot); // split(res.vxa_filename, '.')[0]
        s += "<" + nm + ">";
        s += "<currentTime>" + fmt_double(r[i].current_time) + "</currentTime>";
        s += "<fitness_score>" + fmt_double(r[i].fitness_score) + "</fitness_score>";
        s += "<num_voxel>" + std::to_string(r[i].num_voxel) + "</num_voxel>";
        s += "<num_measured_voxel>" + std::to_string(r[i].num_measured_voxel) + "</num_measured_voxel>";
        s += "<voxSize>" + fmt_double(r[i].vox_size) + "</voxSize>";
        s += "<numClosePairs>" + std::to_string(r[i].num_close_pairs) + "</numClosePairs>";
        s += "<initialCenterOfMass><x>" + fmt_double(r[i].initial_com[0]) + "</x><y>" + fmt_double(r[i].initial_com[1]) + "</y><z>" + fmt_double(r[i].initial_com[2]) + "</z></initialCenterOfMass>";
        s += "<currentCenterOfMass><x>" + fmt_double(r[i].current_com[0]) + "</x><y>" + fmt_double(r[i].current_com[1]) + "</y><z>" + fmt_double(r[i].current_com[2]) + "</z></currentCenterOfMass>";
        s += "<total_distance_of_all_voxels>" + fmt_double(r[i].total_distance_of_all_voxels) + "</total_distance_of_all_voxels>";
        if (vp && vp[i].n_voxels > 0) { // SavePositionOfAllVoxels (vx3_node_worker.cu:122-139): std::to_string(double) is "%f"
            auto triples = [&](const double *p) {
                std::string t;
                char buf[128];
                for (int v = 0; v < vp[i].n_voxels; v++) {
                    snprintf(buf, sizeof(buf), "%f,%f,%f;", p[3 * v], p[3 * v + 1], p[3 * v + 2]);
                    t += buf;
                }
                return t;
            };
            s += "<init_pos>" + triples(vp[i].init_pos) + "</init_pos>";
            s += "<pos>" + triples(vp[i].pos) + "</pos>";
            std::string mats;
            for (int v = 0; v < vp[i].n_voxels; v++) mats += std::to_string(vp[i].mats[v]) + ";";
            s += "<mats>" + mats + "</mats>";
        }
        s += "</" + nm + ">";
    }
    s += "</detail></report>\n";
    fwrite(s.data(), 1, s.size(), f);
    fclose(f);
    return VX3_OK;
}

extern "C" int vx3_worker_run_files(const char *base_vxa, const char *input_dir, const char *const *vxd_files, int n, const char *vxr_path,
                                    const vx3_worker_opts *opts, vx3_result *results_out) {
    if (!base_vxa || n < 0) return VX3_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        vx3_model_set_error("ERROR: No GPU found.");
        return VX3_ERR_NO_DEVICE;
    }
    if (opts && opts->n_devices > 0 && opts->n_devices < ndev) ndev = opts->n_devices;
    const bool verbose = opts && opts->verbose;
    if (verbose) printf("%d GPU found.\n", ndev);
    std::string base_text;
    if (!xml_read_file(base_vxa, &base_text)) {
        vx3_model_set_error(std::string("cannot read ") + base_vxa);
        return VX3_ERR_INVALID;
    }
    // a task without VXD files runs the base VXA alone (convenience; the reference needs at least one VXD)
    std::vector<std::string> files;
    for (int i = 0; i < n; i++) files.push_back(vxd_files[i]);
    const bool base_only = files.empty();
    if (base_only) files.push_back("");
    const int total = (int)files.size();
    std::vector<vx3_result> all(total);
    struct Dump { // per-voxel outputs of a simulation whose VXA sets SavePositionOfAllVoxels
        std::vector<double> init_pos, pos;
        std::vector<int32_t> mats;
    };
    std::vector<Dump> dumps(total);
    std::vector<int> rc(ndev, VX3_OK);
    std::vector<std::string> errs(ndev);
    std::vector<std::thread> threads;
    for (int dev = 0; dev < ndev; dev++) {
        threads.emplace_back([&, dev]() {
            std::vector<int> mine; // sub_batches[i % nDevices] (vx3_node_worker.cu:88-93)
            for (int i = dev; i < total; i += ndev) mine.push_back(i);
            if (mine.empty()) return;
            std::vector<vx3_builder *> builders;
            std::vector<vx3_model_desc> descs;
            auto cleanup = [&]() { for (auto *b : builders) vx3_builder_destroy(b); };
            for (int i : mine) {
                std::string vxd_text, name = base_only ? base_name(base_vxa) : base_name(files[i]);
                if (!base_only && !xml_read_file(join_path(input_dir ? input_dir : "", files[i]), &vxd_text)) {
                    rc[dev] = VX3_ERR_INVALID;
                    errs[dev] = "cannot read " + files[i];
                    cleanup();
                    return;
                }
                vx3_builder *b = vx3_vxa_parse(base_text.c_str(), base_only ? nullptr : vxd_text.c_str(), name.c_str());
                const vx3_model_desc *d = b ? vx3_builder_build(b) : nullptr;
                if (!d) {
                    rc[dev] = VX3_ERR_INVALID;
                    errs[dev] = name + ": " + vx3_model_last_error();
                    if (b) vx3_builder_destroy(b);
                    cleanup();
                    return;
                }
                builders.push_back(b);
                descs.push_back(*d);
            }
            vx3_batch *batch = nullptr;
            int r = vx3_batch_create(dev, descs.data(), (int)descs.size(), &batch);
            if (r == VX3_OK) {
                vx3_run_opts ro;
                memset(&ro, 0, sizeof(ro));
                ro.max_steps = opts ? opts->max_steps : 0;
                ro.emit_history = opts ? opts->emit_history : 1;
                r = vx3_batch_run(batch, &ro, history_to_stdout, nullptr);
            }
            std::vector<vx3_result> res(descs.size());
            if (r == VX3_OK) r = vx3_batch_results(batch, res.data());
            for (size_t k = 0; k < mine.size() && r == VX3_OK; k++) {
                if (!descs[k].opt.save_position_of_all_voxels) continue;
                Dump &dp = dumps[mine[k]];
                const size_t nv = (size_t)descs[k].n_voxels;
                dp.init_pos.resize(3 * nv);
                dp.pos.resize(3 * nv);
                dp.mats.resize(nv);
                r = vx3_batch_positions(batch, (int)k, dp.init_pos.data(), dp.pos.data(), dp.mats.data());
            }
            if (r != VX3_OK) {
                rc[dev] = r;
                errs[dev] = vx3_last_error();
            } else
                for (size_t k = 0; k < mine.size(); k++) all[mine[k]] = res[k];
            if (batch) vx3_batch_destroy(batch);
            cleanup();
        });
    }
    for (auto &t : threads) t.join();
    for (int dev = 0; dev < ndev; dev++)
        if (rc[dev] != VX3_OK) {
            vx3_model_set_error("device " + std::to_string(dev) + ": " + errs[dev]);
            return rc[dev];
        }
    // sortResults: fitness descending, NaN last (the rule of vx3_sort_results), applied to an index so that the per-voxel
    // dumps follow their results
    std::vector<int> order(total);
    for (int i = 0; i < total; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        const double fa = all[a].fitness_score, fb = all[b].fitness_score;
        if (std::isnan(fa)) return false;
        if (std::isnan(fb)) return true;
        return fa > fb;
    });
    std::vector<vx3_result> sorted(total);
    std::vector<vx3_voxel_positions> vps(total);
    bool any_dump = false;
    for (int i = 0; i < total; i++) {
        sorted[i] = all[order[i]];
        const Dump &dp = dumps[order[i]];
        vps[i] = vx3_voxel_positions{(int32_t)dp.mats.size(), 0, dp.init_pos.data(), dp.pos.data(), dp.mats.data()};
        any_dump |= !dp.mats.empty();
    }
    if (results_out) memcpy(results_out, sorted.data(), sizeof(vx3_result) * total);
    if (vxr_path && *vxr_path) return vx3_write_report_positions(vxr_path, input_dir, sorted.data(), any_dump ? vps.data() : nullptr, total);
    return VX3_OK;
}

extern "C" int vx3_worker_run_vxt(const char *vxt_path, const char *vxr_path, const vx3_worker_opts *opts) {
    std::string text, err;
    if (!vxt_path || !xml_read_file(vxt_path, &text)) {
        vx3_model_set_error("Error: input file not found.");
        return VX3_ERR_INVALID;
    }
    std::unique_ptr<XNode> doc = xml_parse(text, &err);
    if (!doc) {
        vx3_model_set_error(err);
        return VX3_ERR_INVALID;
    }
    const XNode *vxa = doc->child("vxa"), *dir = doc->child("input_dir"), *vxd = doc->child("vxd");
    if (!vxa || !dir) {
        vx3_model_set_error("vxt: <vxa> / <input_dir> missing");
        return VX3_ERR_INVALID;
    }
    std::vector<std::string> files;
    if (vxd) for (auto &k : vxd->kids) files.push_back(xml_trim(k->text));
    std::vector<const char *> ptrs;
    for (auto &f : files) ptrs.push_back(f.c_str());
    const std::string base = xml_trim(vxa->text), input_dir = xml_trim(dir->text);
    return vx3_worker_run_files(base.c_str(), input_dir.c_str(), ptrs.data(), (int)ptrs.size(), vxr_path, opts, nullptr);
}
