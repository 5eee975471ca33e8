// gl_mesh_text.cpp -- mesh ingestion behind the C ABI: gemlab's MSH text format (Mesh::read / Mesh::from_text,
// src/mesh/read_text_mesh.rs:36-256, 312-450; README "MSH file format") and its JSON format (Mesh::read_json, src/mesh/mesh.rs:
// 245-256 -- the serde image of `Mesh { ndim, points: [{id, marker, coords}], cells: [{id, marker, kind, points}], marked_edges,
// marked_faces }`, GeoKind as its variant name).  Host code: the parsed mesh is a gl_host_mesh (flat arrays, the layout
// gl_mesh_upload takes); gl_mesh_from_text parses and uploads in one call.  Grammar and error strings are the reference's.
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "gl_internal.hpp"

using namespace gl;

struct gl_host_mesh {
    int ndim = 0;
    std::vector<double> coords;        // npoint x ndim
    std::vector<int32_t> point_markers;
    std::vector<uint8_t> kinds;
    std::vector<int32_t> markers;
    std::vector<uint64_t> conn_off, conn;
    std::vector<int64_t> edges;        // nedge x 3: marker, p1, p2
    std::vector<int64_t> faces;        // nface x 5: marker, p1, p2, p3, p4 (-1 = no fourth point; usize::MAX in the reference)
};

namespace {

thread_local std::string g_text_error;

int text_fail(const char *msg) {
    g_text_error = msg;
    return GL_ERR_MESH;
}

// ---- tokens -----------------------------------------------------------------------------------------------------------------
struct Line {
    std::vector<std::string> f;
};

bool next_fields(const char *&p, const char *end, Line &out) {
    // next non-empty, non-comment line split at whitespace; false at the end of the text
    while (p < end) {
        const char *e = p;
        while (e < end && *e != '\n') e++;
        const char *q = p;
        while (q < e && (*q == ' ' || *q == '\t' || *q == '\r')) q++;
        const bool skip = q == e || *q == '#';
        if (!skip) {
            out.f.clear();
            while (q < e) {
                while (q < e && (*q == ' ' || *q == '\t' || *q == '\r')) q++;
                const char *t = q;
                while (q < e && !(*q == ' ' || *q == '\t' || *q == '\r')) q++;
                if (q > t) out.f.emplace_back(t, q);
            }
        }
        p = e < end ? e + 1 : end;
        if (!skip) return true;
    }
    return false;
}

// Rust's str::parse::<usize>: optional '+', decimal digits only
bool parse_usize(const std::string &s, uint64_t &v) {
    size_t i = 0;
    if (i < s.size() && s[i] == '+') i++;
    if (i >= s.size()) return false;
    unsigned __int128 acc = 0;
    for (; i < s.size(); i++) {
        if (s[i] < '0' || s[i] > '9') return false;
        acc = acc * 10 + (unsigned)(s[i] - '0');
        if (acc > (unsigned __int128)UINT64_MAX) return false;
    }
    v = (uint64_t)acc;
    return true;
}

bool parse_i32(const std::string &s, int32_t &v) {
    size_t i = 0;
    bool neg = false;
    if (i < s.size() && (s[i] == '+' || s[i] == '-')) neg = s[i++] == '-';
    if (i >= s.size()) return false;
    int64_t acc = 0;
    for (; i < s.size(); i++) {
        if (s[i] < '0' || s[i] > '9') return false;
        acc = acc * 10 + (s[i] - '0');
        if (acc > (int64_t)INT32_MAX + 1) return false;
    }
    acc = neg ? -acc : acc;
    if (acc < INT32_MIN || acc > INT32_MAX) return false;
    v = (int32_t)acc;
    return true;
}

bool parse_f64(const std::string &s, double &v) {
    if (s.empty()) return false;
    // Rust's f64::from_str: decimal / exponent forms, "inf", "infinity", "nan" (case-insensitive); no hex, no leading whitespace
    for (char c : s)
        if (c == 'x' || c == 'X' || c == 'p' || c == 'P') return false;
    char *endp = nullptr;
    errno = 0;
    v = std::strtod(s.c_str(), &endp);
    return endp == s.c_str() + s.size();
}

// ---- MSH text -----------------------------------------------------------------------------------------------------------------
int parse_msh(const char *text, uint64_t len, gl_host_mesh &m) {
    const char *p = text, *end = text + len;
    Line ln;
    if (!next_fields(p, end, ln)) return text_fail("file is empty or header is missing");
    static const char *names[5] = {"ndim", "npoint", "ncell", "nmarked_edge", "nmarked_face"};
    uint64_t hdr[5];
    for (int i = 0; i < 5; i++) {
        if ((size_t)i >= ln.f.size()) return text_fail((std::string("cannot read ") + names[i]).c_str());
        if (!parse_usize(ln.f[i], hdr[i])) return text_fail((std::string("cannot parse ") + names[i]).c_str());
    }
    const uint64_t ndim = hdr[0], npoint = hdr[1], ncell = hdr[2], nedge = hdr[3], nface = hdr[4];
    if (ndim < 2 || ndim > 3) return text_fail("space_ndim must be 2 or 3");
    m.ndim = (int)ndim;

    // points: id marker x y [z]
    uint64_t np = 0;
    while (np < npoint && next_fields(p, end, ln)) {
        uint64_t id;
        if (!parse_usize(ln.f[0], id)) return text_fail("cannot parse point id");
        if (id != np) return text_fail("the id and index of points must equal each other");
        if (ln.f.size() < 2) return text_fail("cannot read point marker");
        int32_t marker;
        if (!parse_i32(ln.f[1], marker)) return text_fail("cannot parse point marker");
        static const char *axis[3] = {"x", "y", "z"};
        for (uint64_t j = 0; j < ndim; j++) {
            if (ln.f.size() < 3 + j) return text_fail((std::string("cannot read point ") + axis[j] + " coordinate").c_str());
            double v;
            if (!parse_f64(ln.f[2 + j], v)) return text_fail((std::string("cannot parse point ") + axis[j] + " coordinate").c_str());
            m.coords.push_back(v);
        }
        if (ln.f.size() > 2 + ndim) return text_fail("point data contains extra values");
        m.point_markers.push_back(marker);
        np++;
    }
    if (np != npoint) return text_fail("not all points have been found");

    // cells: id marker kind p0 p1 ..
    uint64_t nc = 0;
    m.conn_off.push_back(0);
    while (nc < ncell && next_fields(p, end, ln)) {
        uint64_t id;
        if (!parse_usize(ln.f[0], id)) return text_fail("cannot parse cell id");
        if (id != nc) return text_fail("the id and index of cells must equal each other");
        if (ln.f.size() < 2) return text_fail("cannot read cell marker");
        int32_t marker;
        if (!parse_i32(ln.f[1], marker)) return text_fail("cannot parse cell marker");
        if (ln.f.size() < 3) return text_fail("cannot read cell kind");
        const int kind = kind_from_name(ln.f[2].c_str());
        if (kind < 0) return text_fail("string representation of GeoKind is incorrect");
        const size_t nn = (size_t)kind_nnode(kind);
        if (ln.f.size() < 3 + nn) return text_fail("cannot read cell point id");
        for (size_t k = 0; k < nn; k++) {
            uint64_t pid;
            if (!parse_usize(ln.f[3 + k], pid)) return text_fail("cannot parse cell point id");
            m.conn.push_back(pid);
        }
        if (ln.f.size() > 3 + nn) return text_fail("cell data contains extra values");
        m.kinds.push_back((uint8_t)kind);
        m.markers.push_back(marker);
        m.conn_off.push_back(m.conn.size());
        nc++;
    }
    if (nc != ncell) return text_fail("not all cells have been found");

    auto marked = [&](uint64_t count, const char *what, int npts_min, int npts_max, std::vector<int64_t> &out, int width) -> int {
        uint64_t n = 0;
        while (n < count && next_fields(p, end, ln)) {
            int32_t marker;
            if (!parse_i32(ln.f[0], marker)) return text_fail((std::string("cannot parse marked ") + what).c_str());
            out.push_back(marker);
            int got = 0;
            for (int j = 0; j < npts_max; j++) {
                if (ln.f.size() < (size_t)(2 + j)) {
                    if (j < npts_min) return text_fail((std::string("cannot read p") + std::to_string(j + 1) + " of marked " + what).c_str());
                    break;
                }
                uint64_t pid;
                if (!parse_usize(ln.f[1 + j], pid)) return text_fail((std::string("cannot parse p") + std::to_string(j + 1) + " of marked " + what).c_str());
                out.push_back((int64_t)pid);
                got++;
            }
            if (ln.f.size() > (size_t)(1 + npts_max)) return text_fail((std::string("marked ") + what + " data contains extra values").c_str());
            for (int j = got; j < width - 1; j++) out.push_back(-1);
            n++;
        }
        if (n != count) return text_fail((std::string("not all marked ") + what + "s have been found").c_str());
        return GL_OK;
    };
    int st = marked(nedge, "edge", 2, 2, m.edges, 3);
    if (st) return st;
    return marked(nface, "face", 3, 4, m.faces, 5);
}

// ---- JSON (just enough for serde_json's image of Mesh) -----------------------------------------------------------------------------
struct JVal {
    enum Type { NUL, BOOL, NUM, STR, ARR, OBJ } type = NUL;
    std::string raw; // number token or string contents
    std::vector<JVal> arr;
    std::vector<std::pair<std::string, JVal>> obj;
    const JVal *get(const char *key) const {
        for (auto &kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
};

struct JParser {
    const char *p, *end;
    bool ok = true;
    void ws() {
        while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++;
    }
    bool lit(const char *s) {
        const size_t n = std::strlen(s);
        if ((size_t)(end - p) < n || std::strncmp(p, s, n) != 0) return false;
        p += n;
        return true;
    }
    std::string str() {
        std::string out;
        if (p >= end || *p != '"') { ok = false; return out; }
        p++;
        while (p < end && *p != '"') {
            if (*p == '\\') {
                p++;
                if (p >= end) { ok = false; return out; }
                switch (*p) {
                case 'n': out += '\n'; break;
                case 't': out += '\t'; break;
                case 'r': out += '\r'; break;
                case 'b': out += '\b'; break;
                case 'f': out += '\f'; break;
                case 'u': if (end - p < 5) { ok = false; return out; } out += '?'; p += 4; break; // not needed for GeoKind names
                default: out += *p;
                }
                p++;
            } else {
                out += *p++;
            }
        }
        if (p >= end) { ok = false; return out; }
        p++;
        return out;
    }
    JVal value(int depth = 0) {
        JVal v;
        ws();
        if (p >= end || depth > 64) { ok = false; return v; }
        if (*p == '{') {
            v.type = JVal::OBJ;
            p++;
            ws();
            if (p < end && *p == '}') { p++; return v; }
            while (ok) {
                ws();
                std::string k = str();
                ws();
                if (!ok || p >= end || *p != ':') { ok = false; break; }
                p++;
                v.obj.emplace_back(std::move(k), value(depth + 1));
                ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == '}') { p++; break; }
                ok = false;
            }
        } else if (*p == '[') {
            v.type = JVal::ARR;
            p++;
            ws();
            if (p < end && *p == ']') { p++; return v; }
            while (ok) {
                v.arr.push_back(value(depth + 1));
                ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == ']') { p++; break; }
                ok = false;
            }
        } else if (*p == '"') {
            v.type = JVal::STR;
            v.raw = str();
        } else if (lit("true") || lit("false")) {
            v.type = JVal::BOOL;
        } else if (lit("null")) {
            v.type = JVal::NUL;
        } else {
            v.type = JVal::NUM;
            const char *t = p;
            while (p < end && (std::strchr("+-0123456789.eE", *p) != nullptr)) p++;
            if (p == t) ok = false;
            v.raw.assign(t, p);
        }
        return v;
    }
};

bool j_usize(const JVal *v, uint64_t &out) { return v && v->type == JVal::NUM && parse_usize(v->raw, out); }
bool j_i32(const JVal *v, int32_t &out) { return v && v->type == JVal::NUM && parse_i32(v->raw, out); }
bool j_f64(const JVal *v, double &out) { return v && v->type == JVal::NUM && parse_f64(v->raw, out); }

// GeoKind serialises as its variant name ("Hex8"); kind_from_name takes the lower-case key ("hex8")
int kind_from_variant(const std::string &name) {
    if (name.empty() || !(name[0] >= 'A' && name[0] <= 'Z')) return -1;
    std::string key = name;
    key[0] = (char)(key[0] - 'A' + 'a');
    const int k = kind_from_name(key.c_str());
    if (k < 0) return -1;
    return k;
}

int parse_json(const char *text, uint64_t len, gl_host_mesh &m) {
    const char *bad = "cannot parse JSON file"; // Mesh::read_json maps every serde error to this string
    JParser jp{text, text + len};
    JVal root = jp.value();
    jp.ws();
    if (!jp.ok || jp.p != jp.end || root.type != JVal::OBJ) return text_fail(bad);
    uint64_t ndim;
    if (!j_usize(root.get("ndim"), ndim) || ndim < 2 || ndim > 3) return text_fail(bad);
    m.ndim = (int)ndim;
    const JVal *pts = root.get("points"), *cells = root.get("cells"), *edges = root.get("marked_edges"), *faces = root.get("marked_faces");
    if (!pts || pts->type != JVal::ARR || !cells || cells->type != JVal::ARR || !edges || edges->type != JVal::ARR || !faces ||
        faces->type != JVal::ARR)
        return text_fail(bad);
    for (size_t i = 0; i < pts->arr.size(); i++) {
        const JVal &pt = pts->arr[i];
        uint64_t id;
        int32_t marker;
        const JVal *c = pt.get("coords");
        if (pt.type != JVal::OBJ || !j_usize(pt.get("id"), id) || !j_i32(pt.get("marker"), marker) || !c || c->type != JVal::ARR ||
            c->arr.size() != ndim)
            return text_fail(bad);
        if (id != i) return text_fail("the id and index of points must equal each other");
        for (auto &x : c->arr) {
            double v;
            if (!j_f64(&x, v)) return text_fail(bad);
            m.coords.push_back(v);
        }
        m.point_markers.push_back(marker);
    }
    m.conn_off.push_back(0);
    for (size_t e = 0; e < cells->arr.size(); e++) {
        const JVal &cl = cells->arr[e];
        uint64_t id;
        int32_t marker;
        const JVal *k = cl.get("kind"), *pp = cl.get("points");
        if (cl.type != JVal::OBJ || !j_usize(cl.get("id"), id) || !j_i32(cl.get("marker"), marker) || !k || k->type != JVal::STR || !pp ||
            pp->type != JVal::ARR)
            return text_fail(bad);
        if (id != e) return text_fail("the id and index of cells must equal each other");
        const int kind = kind_from_variant(k->raw);
        if (kind < 0 || pp->arr.size() != (size_t)kind_nnode(kind)) return text_fail(bad);
        for (auto &x : pp->arr) {
            uint64_t pid;
            if (!j_usize(&x, pid)) return text_fail(bad);
            m.conn.push_back(pid);
        }
        m.kinds.push_back((uint8_t)kind);
        m.markers.push_back(marker);
        m.conn_off.push_back(m.conn.size());
    }
    auto tuples = [&](const JVal *list, size_t width, std::vector<int64_t> &out) -> bool {
        for (auto &t : list->arr) {
            if (t.type != JVal::ARR || t.arr.size() != width) return false;
            int32_t marker;
            if (!j_i32(&t.arr[0], marker)) return false;
            out.push_back(marker);
            for (size_t j = 1; j < width; j++) {
                uint64_t pid;
                if (!j_usize(&t.arr[j], pid)) return false;
                out.push_back(pid == UINT64_MAX ? -1 : (int64_t)pid); // usize::MAX marks the missing fourth point of a triangle
            }
        }
        return true;
    };
    if (!tuples(edges, 3, m.edges) || !tuples(faces, 5, m.faces)) return text_fail(bad);
    return GL_OK;
}

int check_connectivity(const gl_host_mesh &m) {
    const uint64_t npoint = m.point_markers.size();
    for (uint64_t pid : m.conn)
        if (pid >= npoint) return text_fail("cell point id is out of range");
    return GL_OK;
}

} // namespace

extern "C" {

const char *gl_mesh_text_last_error(void) { return g_text_error.c_str(); }

int gl_host_mesh_parse(const char *text, uint64_t len, int format, gl_host_mesh **out) {
    if (!text || !out) return GL_ERR_INVALID_ARG;
    auto *m = new gl_host_mesh();
    int st = format == GL_MESH_FORMAT_JSON ? parse_json(text, len, *m) : (format == GL_MESH_FORMAT_MSH ? parse_msh(text, len, *m) : text_fail("unknown mesh format"));
    if (!st) st = check_connectivity(*m);
    if (st) {
        delete m;
        return st;
    }
    *out = m;
    return GL_OK;
}

int gl_host_mesh_read(const char *path, int format, gl_host_mesh **out) {
    if (!path || !out) return GL_ERR_INVALID_ARG;
    std::FILE *fh = std::fopen(path, "rb");
    if (!fh) return text_fail("cannot open file");
    std::string text;
    char buf[1 << 16];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof(buf), fh)) > 0) text.append(buf, n);
    std::fclose(fh);
    return gl_host_mesh_parse(text.data(), text.size(), format, out);
}

void gl_host_mesh_free(gl_host_mesh *m) { delete m; }
int gl_host_mesh_ndim(const gl_host_mesh *m) { return m ? m->ndim : 0; }
uint64_t gl_host_mesh_npoint(const gl_host_mesh *m) { return m ? m->point_markers.size() : 0; }
uint64_t gl_host_mesh_ncell(const gl_host_mesh *m) { return m ? m->kinds.size() : 0; }
uint64_t gl_host_mesh_nmarked_edge(const gl_host_mesh *m) { return m ? m->edges.size() / 3 : 0; }
uint64_t gl_host_mesh_nmarked_face(const gl_host_mesh *m) { return m ? m->faces.size() / 5 : 0; }
const double *gl_host_mesh_coords(const gl_host_mesh *m) { return m ? m->coords.data() : nullptr; }
const int32_t *gl_host_mesh_point_markers(const gl_host_mesh *m) { return m ? m->point_markers.data() : nullptr; }
const uint8_t *gl_host_mesh_kinds(const gl_host_mesh *m) { return m ? m->kinds.data() : nullptr; }
const int32_t *gl_host_mesh_markers(const gl_host_mesh *m) { return m ? m->markers.data() : nullptr; }
const uint64_t *gl_host_mesh_conn_off(const gl_host_mesh *m) { return m ? m->conn_off.data() : nullptr; }
const uint64_t *gl_host_mesh_conn(const gl_host_mesh *m) { return m ? m->conn.data() : nullptr; }
const int64_t *gl_host_mesh_marked_edges(const gl_host_mesh *m) { return m ? m->edges.data() : nullptr; }
const int64_t *gl_host_mesh_marked_faces(const gl_host_mesh *m) { return m ? m->faces.data() : nullptr; }

int gl_mesh_upload_host(gl_ctx *ctx, const gl_host_mesh *m, gl_mesh **mesh) {
    if (!ctx || !m || !mesh) return GL_ERR_INVALID_ARG;
    return gl_mesh_upload(ctx, m->ndim, m->point_markers.size(), m->coords.data(), m->kinds.size(), m->kinds.data(), m->markers.data(),
                          m->conn_off.data(), m->conn.data(), mesh);
}

int gl_mesh_from_text(gl_ctx *ctx, const char *text, uint64_t len, int format, gl_mesh **mesh) {
    if (!ctx || !mesh) return GL_ERR_INVALID_ARG;
    gl_host_mesh *hm = nullptr;
    int st = gl_host_mesh_parse(text, len, format, &hm);
    if (st) {
        ctx->last_status = st;
        ctx->last_error = g_text_error;
        return st;
    }
    st = gl_mesh_upload_host(ctx, hm, mesh);
    gl_host_mesh_free(hm);
    return st;
}

} // extern "C"
