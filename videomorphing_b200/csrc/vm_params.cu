// vm_params.cu -- project settings reader and writer (host code only).
//
// parse_config_xml(Parameters&, const std::string&) (Algorithm/param_io.h:8) is stale GPUMorph code in the reference
// (SURVEY R2); the live reader is MdiEditor::ReadXmlFile (UI/MdiEditor.cpp:566-749) for the schema written by
// MdiEditor::WriteXmlFile (751-1040):
//   <project> <stage stage=".."/> <videos .../>
//     <parameters> <weight ssim tps ui temp ssimclamp/> <points image1=".." image2=".." connection=".." num=".."/>
//                  <boundary lock=".."/> <debug iternum dropfactor eps startres/> </parameters> </project>
// Point tracks are flat lists of 5-tuples "x y frame keyflag weight " with an all -1 tuple closing each track;
// connections are 4-tuples "ltrack lidx rtrack ridx " with an all -1 tuple closing each group.
// This is a small hand-written attribute scanner (no Qt / libxml2 in the image); behaviour mirrors the Qt reader:
// absent attributes read as 0, fields of absent elements keep their previous (default) values, tokens are split on
// single spaces, the last (empty) token is ignored, and -- like the reference -- the open point list carries over from
// image1 into image2 when image1 does not end with a terminator.
#include "vm_host.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

bool read_file(const char *path, std::string &out) {
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    char buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) out.append(buf, n);
    fclose(f);
    return true;
}

// the text of the first start tag <name ...> at or after `from` ("" if none)
std::string find_tag(const std::string &s, const char *name, size_t from = 0) {
    std::string open = std::string("<") + name;
    size_t p = from;
    while ((p = s.find(open, p)) != std::string::npos) {
        char c = p + open.size() < s.size() ? s[p + open.size()] : '\0';
        if (c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '>' || c == '/') {
            size_t e = p;
            char q = 0;
            for (; e < s.size(); e++) {                       // '>' inside a quoted attribute value does not close the tag
                if (q) { if (s[e] == q) q = 0; }
                else if (s[e] == '"' || s[e] == '\'') q = s[e];
                else if (s[e] == '>') break;
            }
            return s.substr(p, e - p);
        }
        p += open.size();
    }
    return std::string();
}

bool get_attr(const std::string &tag, const char *name, std::string &val) {
    size_t p = 0;
    std::string key = name;
    while ((p = tag.find(key, p)) != std::string::npos) {
        bool left_ok = p > 0 && (tag[p - 1] == ' ' || tag[p - 1] == '\t' || tag[p - 1] == '\n' || tag[p - 1] == '\r');
        size_t q = p + key.size();
        while (q < tag.size() && (tag[q] == ' ' || tag[q] == '\t')) q++;
        if (left_ok && q < tag.size() && tag[q] == '=') {
            q++;
            while (q < tag.size() && (tag[q] == ' ' || tag[q] == '\t')) q++;
            if (q < tag.size() && (tag[q] == '"' || tag[q] == '\'')) {
                char quote = tag[q];
                size_t e = tag.find(quote, q + 1);
                if (e == std::string::npos) return false;
                val = tag.substr(q + 1, e - q - 1);
                return true;
            }
        }
        p += key.size();
    }
    return false;
}
// QString::toInt / toFloat: 0 when the whole string is not a number
int to_int(const std::string &s) {
    if (s.empty()) return 0;
    char *e = nullptr;
    long v = strtol(s.c_str(), &e, 10);
    return (e && *e == '\0') ? (int)v : 0;
}
float to_float(const std::string &s) {
    if (s.empty()) return 0.0f;
    char *e = nullptr;
    float v = strtof(s.c_str(), &e);
    return (e && *e == '\0') ? v : 0.0f;
}
float attr_f(const std::string &tag, const char *name) { std::string v; return get_attr(tag, name, v) ? to_float(v) : 0.0f; }
int attr_i(const std::string &tag, const char *name) { std::string v; return get_attr(tag, name, v) ? to_int(v) : 0; }

std::vector<std::string> split_space(const std::string &s) {     // QString::split(" "), empty parts kept
    std::vector<std::string> out;
    size_t p = 0;
    while (true) {
        size_t e = s.find(' ', p);
        if (e == std::string::npos) { out.push_back(s.substr(p)); break; }
        out.push_back(s.substr(p, e - p));
        p = e + 1;
    }
    return out;
}

template <class T> T *dup_array(const std::vector<T> &v) {
    T *p = static_cast<T *>(malloc(sizeof(T) * (v.empty() ? 1 : v.size())));
    if (p && !v.empty()) memcpy(p, v.data(), sizeof(T) * v.size());
    return p;
}

}  // namespace

using namespace vm;

extern "C" {

int vm_params_parse_xml(const char *path, vm_params *out, vm_tracks *tracks_out) {
    if (!path || !out) { set_error("vm_params_parse_xml: null argument"); return VM_ERR_ARG; }
    std::string s;
    if (!read_file(path, s)) { set_error("vm_params_parse_xml: cannot open %s", path); return VM_ERR_PARSE; }
    if (find_tag(s, "project").empty()) { set_error("vm_params_parse_xml: %s has no <project> element", path); return VM_ERR_PARSE; }
    vm_params_default(out);
    std::string par = find_tag(s, "parameters");
    size_t ppos = par.empty() ? 0 : s.find(par);
    std::vector<std::vector<vm_conp>> lp, rp;
    std::vector<std::vector<vm_connect>> cnt;
    if (!par.empty()) {
        std::string t = find_tag(s, "weight", ppos);
        if (!t.empty()) {                                        // UI/MdiEditor.cpp:645-652
            out->w_ssim = attr_f(t, "ssim"); out->w_tps = attr_f(t, "tps"); out->w_ui = attr_f(t, "ui");
            out->w_temp = attr_f(t, "temp"); out->ssim_clamp = attr_f(t, "ssimclamp");
        }
        t = find_tag(s, "points", ppos);
        if (!t.empty()) {                                        // UI/MdiEditor.cpp:653-716
            std::vector<vm_conp> pt_list;                        // shared by image1 and image2, like the reference
            for (int side = 0; side < 2; side++) {
                std::string v; get_attr(t, side ? "image2" : "image1", v);
                std::vector<std::string> list = split_space(v);
                for (int i = 0; i < (int)list.size() - 1 && i + 4 < (int)list.size(); i += 5) {
                    vm_conp e;
                    e.x = to_int(list[i]); e.y = to_int(list[i + 1]); e.z = to_int(list[i + 2]); e.w = to_int(list[i + 3]);
                    e.weight = to_float(list[i + 4]);
                    if (e.x != -1 || e.y != -1 || e.z != -1 || e.w != -1) pt_list.push_back(e);
                    else { (side ? rp : lp).push_back(pt_list); pt_list.clear(); }
                }
            }
            std::string v; get_attr(t, "connection", v);
            std::vector<std::string> list = split_space(v);
            std::vector<vm_connect> cn_list;
            for (int i = 0; i < (int)list.size() - 1 && i + 3 < (int)list.size(); i += 4) {
                vm_connect e;
                e.li_track = to_int(list[i]); e.li_idx = to_int(list[i + 1]); e.ri_track = to_int(list[i + 2]); e.ri_idx = to_int(list[i + 3]);
                if (e.li_track != -1 || e.li_idx != -1 || e.ri_track != -1 || e.ri_idx != -1) cn_list.push_back(e);
                else { cnt.push_back(cn_list); cn_list.clear(); }
            }
        }
        t = find_tag(s, "boundary", ppos);
        if (!t.empty()) {                                        // UI/MdiEditor.cpp:717-733
            int c = attr_i(t, "lock");
            if (c == 0) out->bcond = VM_BCOND_NONE; else if (c == 1) out->bcond = VM_BCOND_CORNER; else if (c == 2) out->bcond = VM_BCOND_BORDER;
        }
        t = find_tag(s, "debug", ppos);
        if (!t.empty()) {                                        // UI/MdiEditor.cpp:735-741
            out->max_iter = attr_i(t, "iternum"); out->max_iter_drop_factor = attr_f(t, "dropfactor");
            out->eps = attr_f(t, "eps"); out->start_res = attr_i(t, "startres");
        }
    }
    if (tracks_out) {
        memset(tracks_out, 0, sizeof(*tracks_out));
        std::vector<int32_t> ll, rl, gl; std::vector<vm_conp> L, R; std::vector<vm_connect> Cn;
        for (auto &t : lp) { ll.push_back((int32_t)t.size()); L.insert(L.end(), t.begin(), t.end()); }
        for (auto &t : rp) { rl.push_back((int32_t)t.size()); R.insert(R.end(), t.begin(), t.end()); }
        for (auto &g : cnt) { gl.push_back((int32_t)g.size()); Cn.insert(Cn.end(), g.begin(), g.end()); }
        tracks_out->n_left = (int32_t)ll.size(); tracks_out->n_right = (int32_t)rl.size(); tracks_out->n_groups = (int32_t)gl.size();
        tracks_out->left_len = dup_array(ll); tracks_out->right_len = dup_array(rl); tracks_out->group_len = dup_array(gl);
        tracks_out->left = dup_array(L); tracks_out->right = dup_array(R); tracks_out->connects = dup_array(Cn);
    }
    return VM_OK;
}

void vm_tracks_free(vm_tracks *t) {
    if (!t) return;
    free(t->left_len); free(t->right_len); free(t->group_len); free(t->left); free(t->right); free(t->connects);
    memset(t, 0, sizeof(*t));
}


// MdiEditor::WriteXmlFile (UI/MdiEditor.cpp:751-1040), the XML part: <stage>, <videos> (the four attribute values the
// reference always writes; exporting the frames as PNGs / mp4s through avconv is video I/O and out of scope), <parameters>
// with <weight>, <points> (image1 / image2 / connection / num), <boundary>, <debug>.  Numbers are formatted like
// QString::sprintf("%d" / "%f"); every track / connection group is closed by an all -1 tuple; num counts the key points
// (p.w == 1) of both images (MdiEditor.cpp:922,949,978).  QDomDocument::save(out, 4): four spaces per nesting level.
int vm_params_write_xml(const char *path, const vm_params *prm, const vm_tracks *tr, int stage) {
    if (!path || !prm) { set_error("vm_params_write_xml: null argument"); return VM_ERR_ARG; }
    std::string img[2], con;
    int counter = 0;
    char num[64];
    auto app = [&](std::string &dst, const char *fmt, double v, bool is_int) {
        if (is_int) snprintf(num, sizeof(num), fmt, (int)v); else snprintf(num, sizeof(num), fmt, v);
        dst += num;
    };
    if (tr) {
        for (int side = 0; side < 2; side++) {
            const int n = side ? tr->n_right : tr->n_left;
            const int32_t *len = side ? tr->right_len : tr->left_len;
            const vm_conp *pts = side ? tr->right : tr->left;
            size_t o = 0;
            for (int i = 0; i < n; i++) {
                for (int j = 0; j < len[i]; j++, o++) {
                    app(img[side], "%d ", pts[o].x, true); app(img[side], "%d ", pts[o].y, true); app(img[side], "%d ", pts[o].z, true);
                    app(img[side], "%d ", pts[o].w, true); app(img[side], "%f ", pts[o].weight, false);
                    if (pts[o].w == 1) counter++;
                }
                for (int k = 0; k < 4; k++) app(img[side], "%d ", -1, true);
                app(img[side], "%f ", -1.0, false);
            }
        }
        size_t o = 0;
        for (int i = 0; i < tr->n_groups; i++) {
            for (int j = 0; j < tr->group_len[i]; j++, o++) {
                app(con, "%d ", tr->connects[o].li_track, true); app(con, "%d ", tr->connects[o].li_idx, true);
                app(con, "%d ", tr->connects[o].ri_track, true); app(con, "%d ", tr->connects[o].ri_idx, true);
            }
            for (int k = 0; k < 4; k++) app(con, "%d ", -1, true);
        }
    }
    FILE *f = fopen(path, "wb");
    if (!f) { set_error("vm_params_write_xml: cannot open %s for writing", path); return VM_ERR_PARSE; }
    fprintf(f, "<?xml version='1.0'?>\n<project>\n");
    fprintf(f, "    <stage stage=\"%d\"/>\n", stage);
    fprintf(f, "    <videos video1=\"\\video1.mp4\" video2=\"\\video2.mp4\" resample1=\"\\resample1.mp4\" resample2=\"\\resample2.mp4\"/>\n");
    fprintf(f, "    <parameters>\n");
    fprintf(f, "        <weight ssim=\"%f\" tps=\"%f\" ui=\"%f\" temp=\"%f\" ssimclamp=\"%f\"/>\n", prm->w_ssim, prm->w_tps, prm->w_ui, prm->w_temp, prm->ssim_clamp);
    fprintf(f, "        <points image1=\"%s\" image2=\"%s\" connection=\"%s\" num=\"%d\"/>\n", img[0].c_str(), img[1].c_str(), con.c_str(), counter);
    fprintf(f, "        <boundary lock=\"%d\"/>\n", prm->bcond == VM_BCOND_CORNER ? 1 : (prm->bcond == VM_BCOND_BORDER ? 2 : 0));
    fprintf(f, "        <debug iternum=\"%d\" dropfactor=\"%f\" eps=\"%f\" startres=\"%d\"/>\n", prm->max_iter, prm->max_iter_drop_factor, prm->eps, prm->start_res);
    fprintf(f, "    </parameters>\n</project>\n");
    bool ok = ferror(f) == 0;
    ok = (fclose(f) == 0) && ok;
    if (!ok) { set_error("vm_params_write_xml: write to %s failed", path); return VM_ERR_PARSE; }
    return VM_OK;
}

}  // extern "C"
