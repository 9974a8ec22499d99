// Included at the end of vx3_engine.cu (needs the vx3_batch definition).  See vx3_history.h.

namespace vx3 {
std::string HistoryWriter::header(const std::vector<int> &matid, const std::vector<float> &rgba, double vox_size) const {
    // VX3_SimulationManager.cu:40-50
    std::string s = "\n{{{setting}}}<rescale>0.001</rescale>\n";
    char buf[256];
    for (size_t i = 0; i < matid.size(); i++) {
        // the reference prints mat.r/255. etc. as doubles
        snprintf(buf, sizeof(buf), "{{{setting}}}<matcolor><id>%d</id><r>%.2f</r><g>%.2f</g><b>%.2f</b><a>%.2f</a></matcolor>\n", matid[i],
                 (double)rgba[4 * i], (double)rgba[4 * i + 1], (double)rgba[4 * i + 2], (double)rgba[4 * i + 3]);
        s += buf;
    }
    snprintf(buf, sizeof(buf), "\n{{{setting}}}<voxel_size>%f</voxel_size>\n", vox_size);
    s += buf;
    return s;
}
} // namespace vx3

static int history_frame(vx3_batch *b, int sim, long long j, double t, const vx3::HistoryWriter &, std::string &out) {
    const SimC &S = b->simc[sim];
    const vx3_sim_options &o = b->opts[sim];
    const Dev &D = b->D;
    std::vector<SimD> hd;
    int rc = fetch_simd(b, hd);
    if (rc) return rc;
    const int nv = S.nvox, nl = hd[sim].link_cnt;
    std::vector<double> pose;
    std::vector<int32_t> vflags, vlinks, lstate, lmat;
    std::vector<float> tempe;
    std::vector<int2> lends;
    std::vector<float4> lstrain;
    if ((rc = d2h(b, pose, D.pose, 8 * (size_t)S.voff, 8 * (size_t)nv))) return rc;
    if ((rc = d2h(b, vflags, D.vflags, S.voff, nv))) return rc;
    if ((rc = d2h(b, tempe, D.tempe, S.voff, nv))) return rc;
    if ((rc = d2h(b, vlinks, D.vlinks, 6 * (size_t)S.voff, 6 * (size_t)nv))) return rc;
    if ((rc = d2h(b, lends, D.lends, S.loff, nl))) return rc;
    if ((rc = d2h(b, lstate, D.lstate, S.loff, nl))) return rc;
    if ((rc = d2h(b, lmat, D.lmat, S.loff, nl))) return rc;
    if ((rc = d2h(b, lstrain, D.lstrain, S.loff, nl))) return rc;
    std::vector<double> sig;
    if (D.sig && b->simc[sim].enable_signals && (rc = d2h(b, sig, D.sig, 6 * (size_t)S.voff, 6 * (size_t)nv))) return rc;
    CK(cudaStreamSynchronize(b->stream));
    const double vs = 1 / 0.001;
    char buf[512];
    out.clear();
    const std::vector<vx3_voxel_material> &vm = b->h_vmats[sim];
    if (o.record_voxel) {
        snprintf(buf, sizeof(buf), "<<<Step%d Time:%f>>>", (int)j, t);
        out += buf;
        for (int i = 0; i < nv; i++) { // model order; vd / li = position in the batch's storage order
            const size_t vd = (size_t)(b->vdev((size_t)S.voff + i) - S.voff);
            if (vflags[vd] & (VX3_VOX_SURFACE | VXF_REMOVED)) continue; // interior (bit named SURFACE) or removed
            const double *p = &pose[8 * vd];
            const vx3_voxel_material &m = vm[b->vmat_local[sim][i]];
            const Q4 q(p[3], p[4], p[5], p[6]);
            snprintf(buf, sizeof(buf), "%.1f,%.1f,%.1f,", p[0] * vs, p[1] * vs, p[2] * vs);
            out += buf;
            snprintf(buf, sizeof(buf), "%.1f,%.2f,%.2f,%.2f,", q.Angle() * 57.29577951308232, q.x, q.y, q.z);
            out += buf;
            // cornerOffset(NNN), cornerOffset(PPP): VX3_Voxel.cu:147-159 (Vec3D<float> results)
            float corner[2][3];
            for (int c = 0; c < 2; c++) {
                const bool posLink = c == 1;
                for (int a = 0; a < 3; a++) {
                    double strain = posLink ? 1.0 : -1.0;
                    const int g = vlinks[6 * vd + 2 * a + (posLink ? 0 : 1)];
                    if (g >= 0) {
                        const int li = g - S.loff;
                        const LinkMatC &lm = b->h_lmat_tab[lmat[li]];
                        const bool failed = lm.epsilonFail != -1.0f && lstrain[li].y > lm.epsilonFail;
                        if (!failed) {
                            const float En = vm[b->vmat_local[sim][b->vext((size_t)lends[li].x) - S.voff]].E, Ep = vm[b->vmat_local[sim][b->vext((size_t)lends[li].y) - S.voff]].E;
                            const float ratio = Ep / En, st = lstrain[li].x; // strainRatio, strain
                            const float ax = posLink ? 2.0f * st * ratio / (1.0f + ratio) : 2.0f * st / (1.0f + ratio);
                            strain = (1 + ax) * (posLink ? 1 : -1);
                        }
                    }
                    const double base = (m.nomSize * m.extScale[a]) * (1 + tempe[vd] * m.alphaCTE);
                    corner[c][a] = (float)((0.5 * base) * strain);
                }
            }
            snprintf(buf, sizeof(buf), "%.1f,%.1f,%.1f,%.1f,%.1f,%.1f,", corner[0][0] * vs, corner[0][1] * vs, corner[0][2] * vs, corner[1][0] * vs,
                     corner[1][1] * vs, corner[1][2] * vs);
            out += buf;
            snprintf(buf, sizeof(buf), "%d,", m.matid);
            out += buf;
            snprintf(buf, sizeof(buf), "%.1f,", sig.empty() ? 0.0 : sig[6 * vd]); // localSignal (VX3_SimulationManager.cu:88)
            out += buf;
            out += ";";
        }
        out += "<<<>>>";
    }
    if (o.record_link) {
        snprintf(buf, sizeof(buf), "|[[[%d]]]", (int)j);
        out += buf;
        for (int ie = 0; ie < nl; ie++) {
            const size_t i = (size_t)(b->ldev((size_t)S.loff + ie) - S.loff);
            if (lstate[i] & (LKS_REMOVED | LKS_DETACHED)) continue;
            const double *p1 = &pose[8 * (size_t)(lends[i].y - S.voff)], *p2 = &pose[8 * (size_t)(lends[i].x - S.voff)];
            snprintf(buf, sizeof(buf), "%.4f,%.4f,%.4f,%.4f,%.4f,%.4f,;", p1[0], p1[1], p1[2], p2[0], p2[1], p2[2]);
            out += buf;
        }
        out += "[[[]]]";
    }
    out += "\n";
    return VX3_OK;
}
