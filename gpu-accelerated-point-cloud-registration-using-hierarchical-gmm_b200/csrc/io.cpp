// io.cpp -- cloud ingestion, the callers' step in front of the hot path.  Host code only: no CUDA call.
//
//   hgmm_io_read_ply   ASCII PLY vertices with the semantics of either reference reader, or header-driven:
//     HGMM_PLY_HEADER      parse the header (`format ascii`, `element vertex N` with x y z as its first three properties,
//                          `end_header`), read N vertex lines
//     HGMM_PLY_VIEWER_FIT  readData, src/c++/main.cpp:45-79: skip the first 24 non-empty lines, then take lines while they
//                          have exactly three tokens (the range_grid lines that follow the vertices have one or two);
//                          the viewer's two hard-coded similarity transforms (main.cpp:30-38,68,75) are NOT applied
//     HGMM_PLY_VIEWER_REG  readPointCloud, src/c++/main_reg.cpp:106-161: 17 lines, the vertex count is the third token of
//                          line 18, 6 more lines, then that many vertex lines (first three tokens)
//   hgmm_io_read_pcd   PCD v0.7 with FIELDS x y z first, SIZE 4, TYPE F: `DATA ascii` or `DATA binary` (the Waymo sweeps of
//                          src/python/{hgmm,gmmreg_gpu}/waymo*.pcd); `binary_compressed` is refused
// Both fill out_xyz ([N,3] packed float32, the layout hgmm_set_points takes) up to `capacity` points and always report the
// number of points in the file, so a caller can size its buffer with a first call (out_xyz = NULL, capacity = 0).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/hgmm.h"

namespace {

// getline that accepts \n, \r\n and \r line ends (utilityCore::safeGetline, src/c++/common/utilities.cpp)
bool safe_getline(std::istream& is, std::string& line) {
    line.clear();
    std::streambuf* sb = is.rdbuf();
    bool any = false;
    for (;;) {
        const int c = sb->sbumpc();
        if (c == EOF) {
            is.setstate(std::ios::eofbit);
            return any;
        }
        any = true;
        if (c == '\n') return true;
        if (c == '\r') {
            if (sb->sgetc() == '\n') sb->sbumpc();
            return true;
        }
        line.push_back((char)c);
    }
}

void tokenize(const std::string& line, std::vector<std::string>& out) {
    out.clear();
    std::istringstream ss(line);
    std::string t;
    while (ss >> t) out.push_back(t);
}

inline void put(float* out, int64_t cap, int64_t i, const std::vector<std::string>& tok) {
    if (out && i < cap) {
        out[3 * i] = (float)atof(tok[0].c_str());
        out[3 * i + 1] = (float)atof(tok[1].c_str());
        out[3 * i + 2] = (float)atof(tok[2].c_str());
    }
}

}  // namespace

extern "C" {

static int read_ply_impl(const char* path, int32_t mode, float* out_xyz, int64_t capacity, int64_t* out_n);
int hgmm_io_read_ply(const char* path, int32_t mode, float* out_xyz, int64_t capacity, int64_t* out_n) {
    if (!path || !out_n || capacity < 0 || (capacity > 0 && !out_xyz)) return HGMM_ERR_INVALID;
    if (mode != HGMM_PLY_HEADER && mode != HGMM_PLY_VIEWER_FIT && mode != HGMM_PLY_VIEWER_REG) return HGMM_ERR_INVALID;
    *out_n = 0;
    try {
        return read_ply_impl(path, mode, out_xyz, capacity, out_n);
    } catch (...) {
        *out_n = 0;
        return HGMM_ERR_IO;
    }
}
static int read_ply_impl(const char* path, int32_t mode, float* out_xyz, int64_t capacity, int64_t* out_n) {
    std::ifstream in(path, std::ios::binary);
    if (!in.is_open()) return HGMM_ERR_IO;
    std::string line;
    std::vector<std::string> tok;
    int64_t n = 0;
    if (mode == HGMM_PLY_VIEWER_FIT) {
        int count = 0;
        while (in.good()) {
            if (!safe_getline(in, line) && line.empty()) break;
            if (line.empty()) continue;
            ++count;
            if (count < 25) continue;
            tokenize(line, tok);
            if (tok.size() != 3) break;
            put(out_xyz, capacity, n, tok);
            ++n;
        }
        *out_n = n;
        return HGMM_OK;
    }
    int64_t want = -1;
    if (mode == HGMM_PLY_VIEWER_REG) {
        for (int i = 0; i < 17; ++i)
            if (!safe_getline(in, line)) return HGMM_ERR_IO;
        if (!safe_getline(in, line)) return HGMM_ERR_IO;
        tokenize(line, tok);
        if (tok.size() < 3) return HGMM_ERR_IO;
        want = atoll(tok[2].c_str());
        for (int i = 0; i < 6; ++i)
            if (!safe_getline(in, line)) return HGMM_ERR_IO;
    } else {
        if (!safe_getline(in, line)) return HGMM_ERR_IO;
        tokenize(line, tok);
        if (tok.empty() || tok[0] != "ply") return HGMM_ERR_IO;
        bool ascii = false, in_vertex = false, done = false;
        int vprops = 0;
        while (!done && safe_getline(in, line)) {
            tokenize(line, tok);
            if (tok.empty()) continue;
            if (tok[0] == "format") {
                ascii = tok.size() >= 2 && tok[1] == "ascii";
            } else if (tok[0] == "element") {
                in_vertex = tok.size() >= 3 && tok[1] == "vertex";
                if (in_vertex) {
                    if (want >= 0) return HGMM_ERR_IO;       // two vertex elements
                    want = atoll(tok[2].c_str());
                } else if (want < 0) {
                    return HGMM_ERR_IO;                      // another element's data would precede the vertices
                }
            } else if (tok[0] == "property" && in_vertex) {
                static const char* names[3] = {"x", "y", "z"};
                if (vprops < 3 && (tok.size() < 3 || tok.back() != names[vprops])) return HGMM_ERR_IO;
                ++vprops;
            } else if (tok[0] == "end_header") {
                done = true;
            }
        }
        if (!done || !ascii || want < 0 || vprops < 3) return HGMM_ERR_IO;
    }
    if (want < 0) return HGMM_ERR_IO;
    for (int64_t i = 0; i < want; ++i) {
        if (!safe_getline(in, line)) return HGMM_ERR_IO;     // truncated file
        tokenize(line, tok);
        if (tok.size() < 3) return HGMM_ERR_IO;
        put(out_xyz, capacity, i, tok);
    }
    *out_n = want;
    return HGMM_OK;
}

static int read_pcd_impl(const char* path, float* out_xyz, int64_t capacity, int64_t* out_n);
int hgmm_io_read_pcd(const char* path, float* out_xyz, int64_t capacity, int64_t* out_n) {
    if (!path || !out_n || capacity < 0 || (capacity > 0 && !out_xyz)) return HGMM_ERR_INVALID;
    *out_n = 0;
    try {                                    // no exception (bad_alloc on a hostile header, stream failures) crosses the C ABI
        return read_pcd_impl(path, out_xyz, capacity, out_n);
    } catch (...) {
        *out_n = 0;
        return HGMM_ERR_IO;
    }
}
static int read_pcd_impl(const char* path, float* out_xyz, int64_t capacity, int64_t* out_n) {
    std::ifstream in(path, std::ios::binary);
    if (!in.is_open()) return HGMM_ERR_IO;
    std::string line;
    std::vector<std::string> tok;
    std::vector<std::string> fields;
    std::vector<int> sizes, counts;
    std::vector<std::string> types;
    int64_t points = -1, width = -1, height = 1;
    std::string data;
    while (data.empty() && safe_getline(in, line)) {
        tokenize(line, tok);
        if (tok.empty() || tok[0][0] == '#') continue;
        if (tok[0] == "FIELDS") fields.assign(tok.begin() + 1, tok.end());
        else if (tok[0] == "SIZE") for (size_t i = 1; i < tok.size(); ++i) sizes.push_back(atoi(tok[i].c_str()));
        else if (tok[0] == "TYPE") types.assign(tok.begin() + 1, tok.end());
        else if (tok[0] == "COUNT") for (size_t i = 1; i < tok.size(); ++i) counts.push_back(atoi(tok[i].c_str()));
        else if (tok[0] == "WIDTH" && tok.size() > 1) width = atoll(tok[1].c_str());
        else if (tok[0] == "HEIGHT" && tok.size() > 1) height = atoll(tok[1].c_str());
        else if (tok[0] == "POINTS" && tok.size() > 1) points = atoll(tok[1].c_str());
        else if (tok[0] == "DATA" && tok.size() > 1) data = tok[1];
    }
    if (points < 0 && width >= 0) points = width * height;
    if (data.empty() || points < 0 || fields.size() < 3 || fields[0] != "x" || fields[1] != "y" || fields[2] != "z") return HGMM_ERR_IO;
    if (sizes.size() != fields.size() || types.size() != fields.size()) return HGMM_ERR_IO;
    if (!counts.empty() && counts.size() != fields.size()) return HGMM_ERR_IO;           // COUNT must cover every field
    size_t row_bytes = 0;
    for (size_t i = 0; i < fields.size(); ++i) {                                          // untrusted header: every SIZE / COUNT
        const int c = counts.empty() ? 1 : counts[i];                                     // positive, row size bounded
        if (sizes[i] <= 0 || sizes[i] > 8 || c <= 0 || c > 4096) return HGMM_ERR_IO;
        row_bytes += (size_t)sizes[i] * (size_t)c;
        if (row_bytes > (size_t)1 << 20) return HGMM_ERR_IO;
    }
    for (int i = 0; i < 3; ++i)
        if (sizes[i] != 4 || types[i] != "F" || (!counts.empty() && counts[i] != 1)) return HGMM_ERR_IO;
    if (data == "ascii") {
        for (int64_t i = 0; i < points; ++i) {
            if (!safe_getline(in, line)) return HGMM_ERR_IO;
            tokenize(line, tok);
            if (tok.size() < 3) return HGMM_ERR_IO;
            put(out_xyz, capacity, i, tok);
        }
    } else if (data == "binary") {
        const size_t stride = row_bytes;
        std::vector<char> row(stride);
        for (int64_t i = 0; i < points; ++i) {
            in.read(row.data(), (std::streamsize)stride);
            if ((size_t)in.gcount() != stride) return HGMM_ERR_IO;
            if (out_xyz && i < capacity) memcpy(out_xyz + 3 * i, row.data(), 12);      // little-endian float32 x y z lead the row
        }
    } else {
        return HGMM_ERR_IO;                                  // binary_compressed (LZF): not supported
    }
    *out_n = points;
    return HGMM_OK;
}

}  // extern "C"
