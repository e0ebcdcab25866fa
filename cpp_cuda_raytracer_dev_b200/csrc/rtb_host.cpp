// rtb_host.cpp -- mesh load, tree build, camera basis, transform recurrence (host, fp32, no FMA).
#include "rtb_host.hpp"

#include <algorithm>
#include <atomic>
#include <cctype>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

namespace rtb {

// =================================================================================================
// PLY
// =================================================================================================
namespace {

struct FileBytes {
    std::vector<char> data;
    bool read(const char* path) {
        FILE* f = std::fopen(path, "rb");
        if (!f) return false;
        std::fseek(f, 0, SEEK_END);
        long sz = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        data.resize((size_t)sz + 1);
        size_t got = std::fread(data.data(), 1, (size_t)sz, f);
        std::fclose(f);
        data[got] = 0;
        data.resize(got + 1);
        return got == (size_t)sz;
    }
};

inline void emit_triangle(std::vector<float>& out, const float* a, const float* b, const float* c) {
    out.insert(out.end(), a, a + 3);
    out.insert(out.end(), b, b + 3);
    out.insert(out.end(), c, c + 3);
}

// face rules of read_ply.cpp:70-149
inline bool emit_face(std::vector<float>& out, const std::vector<float>& verts, int count, const long* idx) {
    const long nv = (long)(verts.size() / 3);
    for (int k = 0; k < count; k++)
        if (idx[k] < 0 || idx[k] >= nv) return false;
    const float* base = verts.data();
    if (count == 3) {
        emit_triangle(out, base + 3 * idx[2], base + 3 * idx[0], base + 3 * idx[1]);  // stored (c,a,b)
    } else {
        emit_triangle(out, base + 3 * idx[0], base + 3 * idx[1], base + 3 * idx[2]);  // (A,B,C)
        emit_triangle(out, base + 3 * idx[0], base + 3 * idx[2], base + 3 * idx[3]);  // (A,C,D)
    }
    return true;
}

}  // namespace

// ---- conforming PLY reader (mode -1) ---------------------------------------------------------------
// The reference loader (read_ply.cpp:13-153) is a text scanner for four fixed column layouts.  Mode -1 is the
// header-driven reader SURVEY.md section 8(f) item 2 asks for: `format ascii | binary_little_endian |
// binary_big_endian`, any element order, any scalar property types, list properties, extra elements and
// properties skipped by their declared sizes.  Vertices are taken from the properties named x, y, z; faces from the
// list property `vertex_indices` / `vertex_index` (else the first list property of `face`).  The triangle order
// rules stay the reference's (read_ply.cpp:92-148): a 3-gon (a,b,c) is stored (c,a,b), a 4-gon becomes
// (a,b,c),(a,c,d); larger polygons continue that fan: (a,v[k],v[k+1]).
namespace {

enum PlyType { kI8, kU8, kI16, kU16, kI32, kU32, kF32, kF64, kBadType };
const int kPlySize[] = {1, 1, 2, 2, 4, 4, 4, 8, 0};

PlyType ply_type(const std::string& t) {
    if (t == "char" || t == "int8") return kI8;
    if (t == "uchar" || t == "uint8") return kU8;
    if (t == "short" || t == "int16") return kI16;
    if (t == "ushort" || t == "uint16") return kU16;
    if (t == "int" || t == "int32") return kI32;
    if (t == "uint" || t == "uint32") return kU32;
    if (t == "float" || t == "float32") return kF32;
    if (t == "double" || t == "float64") return kF64;
    return kBadType;
}

struct PlyProperty {
    std::string name;
    bool is_list = false;
    PlyType type = kBadType, count_type = kBadType;
};
struct PlyElement {
    std::string name;
    long count = 0;
    std::vector<PlyProperty> props;
};

std::vector<std::string> split_ws(const std::string& line) {
    std::vector<std::string> out;
    size_t i = 0;
    while (i < line.size()) {
        while (i < line.size() && std::isspace((unsigned char)line[i])) i++;
        size_t j = i;
        while (j < line.size() && !std::isspace((unsigned char)line[j])) j++;
        if (j > i) out.emplace_back(line, i, j - i);
        i = j;
    }
    return out;
}

// one scalar from the data section, as a double (exact for every PLY type except that float64 -> float
// narrowing happens at the caller, once)
struct PlyCursor {
    const char* p;
    const char* end;
    int format;  // 0 ascii, 1 little endian, 2 big endian
    bool ok = true;
    double next(PlyType t) {
        if (format == 0) {
            char* q = nullptr;
            double v;
            if (t == kF32) v = (double)std::strtof(p, &q);  // one correctly rounded conversion, like `istream >> float`
            else if (t == kF64) v = std::strtod(p, &q);
            else v = (double)std::strtoll(p, &q, 10);
            if (q == p || q > end) { ok = false; return 0.0; }
            p = q;
            return v;
        }
        const int n = kPlySize[t];
        if (end - p < n) { ok = false; return 0.0; }
        unsigned char b[8];
        for (int k = 0; k < n; k++) b[k] = (unsigned char)(format == 1 ? p[k] : p[n - 1 - k]);  // -> little endian
        p += n;
        switch (t) {
            case kI8: return (double)(int8_t)b[0];
            case kU8: return (double)b[0];
            case kI16: { int16_t v; std::memcpy(&v, b, 2); return (double)v; }
            case kU16: { uint16_t v; std::memcpy(&v, b, 2); return (double)v; }
            case kI32: { int32_t v; std::memcpy(&v, b, 4); return (double)v; }
            case kU32: { uint32_t v; std::memcpy(&v, b, 4); return (double)v; }
            case kF32: { float v; std::memcpy(&v, b, 4); return (double)v; }
            case kF64: { double v; std::memcpy(&v, b, 8); return v; }
            default: ok = false; return 0.0;
        }
    }
};

std::string load_ply_conforming(const FileBytes& fb, std::vector<float>& points9) {
    const char* p = fb.data.data();
    const char* end = p + fb.data.size() - 1;
    std::vector<PlyElement> elements;
    int format = -1;
    bool header_done = false, first = true;
    while (p < end) {
        const char* nl = (const char*)std::memchr(p, '\n', (size_t)(end - p));
        std::string line(p, nl ? nl : end);
        p = nl ? nl + 1 : end;
        const std::vector<std::string> w = split_ws(line);
        if (first) {
            if (w.empty() || w[0] != "ply") return "read_ply: not a PLY file";
            first = false;
            continue;
        }
        if (w.empty() || w[0] == "comment" || w[0] == "obj_info") continue;
        if (w[0] == "end_header") { header_done = true; break; }
        if (w[0] == "format" && w.size() >= 2) {
            format = w[1] == "ascii" ? 0 : w[1] == "binary_little_endian" ? 1 : w[1] == "binary_big_endian" ? 2 : -1;
            if (format < 0) return "read_ply: unknown format " + w[1];
        } else if (w[0] == "element" && w.size() >= 3) {
            PlyElement e;
            e.name = w[1];
            e.count = std::atol(w[2].c_str());
            if (e.count < 0) return "read_ply: negative element count";
            // every element occupies at least one byte of the body (ascii: its line break), whatever the format
            if (e.count > (long)fb.data.size()) return "read_ply: element count exceeds the file size";
            elements.push_back(e);
        } else if (w[0] == "property") {
            if (elements.empty()) return "read_ply: property before any element";
            PlyProperty pr;
            if (w.size() >= 5 && w[1] == "list") {
                pr.is_list = true; pr.count_type = ply_type(w[2]); pr.type = ply_type(w[3]); pr.name = w[4];
                if (pr.count_type == kBadType || pr.count_type == kF32 || pr.count_type == kF64) return "read_ply: bad list count type";
            } else if (w.size() >= 3) {
                pr.type = ply_type(w[1]); pr.name = w[2];
            } else {
                return "read_ply: malformed property line";
            }
            if (pr.type == kBadType) return "read_ply: unknown property type in `" + line + "`";
            elements.back().props.push_back(pr);
        } else {
            return "read_ply: unknown header line `" + line + "`";
        }
    }
    if (!header_done) return "read_ply: no end_header";
    if (format < 0) return "read_ply: header has no format line";

    PlyCursor cur{p, end, format};
    std::vector<float> verts;
    std::vector<long> face_idx;      // all polygons' indices back to back
    std::vector<int> face_count;     // vertices per polygon
    bool have_vertex = false, have_face = false;
    for (const PlyElement& e : elements) {
        if (e.name == "vertex") {
            int ix = -1, iy = -1, iz = -1;
            for (size_t k = 0; k < e.props.size(); k++) {
                if (e.props[k].is_list) continue;
                if (e.props[k].name == "x") ix = (int)k;
                if (e.props[k].name == "y") iy = (int)k;
                if (e.props[k].name == "z") iz = (int)k;
            }
            if (ix < 0 || iy < 0 || iz < 0) return "read_ply: vertex element needs scalar properties x, y, z";
            verts.resize((size_t)e.count * 3);
            have_vertex = true;
            for (long i = 0; i < e.count; i++) {
                for (size_t k = 0; k < e.props.size(); k++) {
                    const PlyProperty& pr = e.props[k];
                    if (pr.is_list) {
                        const long n = (long)cur.next(pr.count_type);
                        for (long j = 0; j < n && cur.ok; j++) cur.next(pr.type);
                    } else {
                        const double v = cur.next(pr.type);
                        if ((int)k == ix) verts[3 * (size_t)i] = (float)v;
                        else if ((int)k == iy) verts[3 * (size_t)i + 1] = (float)v;
                        else if ((int)k == iz) verts[3 * (size_t)i + 2] = (float)v;
                    }
                }
                if (!cur.ok) return "read_ply: truncated or malformed vertex data";
            }
        } else if (e.name == "face") {
            int il = -1;
            for (size_t k = 0; k < e.props.size(); k++)
                if (e.props[k].is_list && (e.props[k].name == "vertex_indices" || e.props[k].name == "vertex_index")) il = (int)k;
            for (size_t k = 0; k < e.props.size() && il < 0; k++)
                if (e.props[k].is_list) il = (int)k;
            if (il < 0) return "read_ply: face element has no list property";
            have_face = true;
            face_count.reserve((size_t)e.count);
            face_idx.reserve((size_t)e.count * 3);
            for (long i = 0; i < e.count; i++) {
                for (size_t k = 0; k < e.props.size(); k++) {
                    const PlyProperty& pr = e.props[k];
                    if (pr.is_list) {
                        const long n = (long)cur.next(pr.count_type);
                        if (!cur.ok || n < 0 || n > (1 << 20)) return "read_ply: truncated or malformed face data";
                        if ((int)k == il) face_count.push_back((int)n);
                        for (long j = 0; j < n && cur.ok; j++) {
                            const double v = cur.next(pr.type);
                            if ((int)k == il) face_idx.push_back((long)v);
                        }
                    } else {
                        cur.next(pr.type);
                    }
                }
                if (!cur.ok) return "read_ply: truncated or malformed face data";
            }
        } else {  // an element the path does not use: step over it by its declared layout
            for (long i = 0; i < e.count; i++) {
                for (const PlyProperty& pr : e.props) {
                    if (pr.is_list) {
                        const long n = (long)cur.next(pr.count_type);
                        for (long j = 0; j < n && cur.ok; j++) cur.next(pr.type);
                    } else {
                        cur.next(pr.type);
                    }
                }
                if (!cur.ok) return "read_ply: truncated or malformed data in element " + e.name;
            }
        }
    }
    if (!have_vertex || !have_face) return "read_ply: header has no vertex/face counts";
    if (verts.empty() || face_count.empty()) return "read_ply: header has no vertex/face counts";
    const long nv = (long)(verts.size() / 3);
    const float* base = verts.data();
    size_t at = 0;
    for (int n : face_count) {
        const long* idx = face_idx.data() + at;
        at += (size_t)n;
        if (n < 3) return "read_ply: a face needs at least three vertices";
        for (int k = 0; k < n; k++)
            if (idx[k] < 0 || idx[k] >= nv) return "read_ply: vertex index out of range";
        if (n == 3) {
            emit_triangle(points9, base + 3 * idx[2], base + 3 * idx[0], base + 3 * idx[1]);  // stored (c,a,b), read_ply.cpp:138-148
        } else {
            for (int k = 1; k + 1 < n; k++) emit_triangle(points9, base + 3 * idx[0], base + 3 * idx[k], base + 3 * idx[k + 1]);
        }
    }
    return "";
}

}  // namespace

std::string load_ply(const char* path, int mode, std::vector<float>& points9) {
    points9.clear();
    if (mode < -1 || mode > 2) return "read_ply: unsupported mode (0, 1, 2 or -1)";
    FileBytes fb;
    if (!fb.read(path)) return std::string("read_ply: cannot read ") + path;
    if (mode == -1) return load_ply_conforming(fb, points9);
    const char* p = fb.data.data();
    const char* end = p + fb.data.size() - 1;
    long num_vert = 0, num_face = 0;
    bool binary = false, header_done = false;
    // header: only `element vertex|face <n>` matter to the reference (read_ply.cpp:19-44)
    while (p < end) {
        const char* nl = (const char*)std::memchr(p, '\n', (size_t)(end - p));
        const char* line_end = nl ? nl : end;
        std::string line(p, line_end);
        while (!line.empty() && (line.back() == '\r' || line.back() == ' ')) line.pop_back();
        p = nl ? nl + 1 : end;
        if (line == "end_header") { header_done = true; break; }
        if (line.rfind("format binary", 0) == 0) binary = true;
        if (line.rfind("element vertex ", 0) == 0) num_vert = std::atol(line.c_str() + 15);
        else if (line.rfind("element face ", 0) == 0) num_face = std::atol(line.c_str() + 13);
    }
    if (!header_done) return "read_ply: no end_header";
    if (num_vert <= 0 || num_face <= 0) return "read_ply: header has no vertex/face counts";
    if (binary) return "read_ply: binary PLY needs mode -1 (the reference reader only parses text)";
    if (num_vert > (long)(end - p) || num_face > (long)(end - p)) return "read_ply: element count exceeds the file size";
    std::vector<float> verts((size_t)num_vert * 3);
    points9.reserve((size_t)num_face * 9);
    const int columns = mode == 1 ? 5 : mode == 2 ? 6 : 3;
    char* cur = const_cast<char*>(p);
    char* next = nullptr;
    for (long i = 0; i < num_vert; i++) {
        for (int c = 0; c < columns; c++) {
            float v = std::strtof(cur, &next);  // `istream >> float`: one correctly rounded conversion
            if (next == cur) return "read_ply: bad vertex record";
            cur = next;
            if (c < 3) verts[3 * (size_t)i + c] = v;
        }
    }
    for (long f = 0; f < num_face; f++) {
        long count = std::strtol(cur, &next, 10);
        if (next == cur) return "read_ply: bad face record";
        cur = next;
        if (count != 3 && count != 4) return "read_ply: only triangles and quads are supported";
        long idx[4] = {0, 0, 0, 0};
        for (long k = 0; k < count; k++) {
            idx[k] = std::strtol(cur, &next, 10);
            if (next == cur) return "read_ply: bad face record";
            cur = next;
        }
        if (!emit_face(points9, verts, (int)count, idx)) return "read_ply: vertex index out of range";
    }
    return "";
}

std::string save_ply(const char* path, const float* points9, uint32_t num_tri) {
    FILE* f = std::fopen(path, "wb");
    if (!f) return std::string("write_ply: cannot open ") + path;
    std::fprintf(f, "ply\nformat ascii 1.0\nelement vertex %u\nproperty float x\nproperty float y\nproperty float z\n"
                    "element face %u\nproperty list uchar int vertex_indices\nend_header\n", num_tri * 3, num_tri);
    for (uint64_t i = 0; i < (uint64_t)num_tri * 3; i++)
        std::fprintf(f, "%.9g %.9g %.9g\n", points9[3 * i], points9[3 * i + 1], points9[3 * i + 2]);
    // the loader stores face (a,b,c) as (c,a,b); writing (v1,v2,v0) restores (v0,v1,v2)
    for (uint64_t t = 0; t < num_tri; t++)
        std::fprintf(f, "3 %llu %llu %llu\n", (unsigned long long)(3 * t + 1), (unsigned long long)(3 * t + 2), (unsigned long long)(3 * t));
    std::fclose(f);
    return "";
}

// =================================================================================================
// frame and tree files
// =================================================================================================
namespace {
uint32_t crc32_update(uint32_t crc, const unsigned char* p, size_t n) {
    static uint32_t table[256];
    static bool ready = false;
    if (!ready) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        ready = true;
    }
    for (size_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
    return crc;
}
void put_be32(std::vector<unsigned char>& v, uint32_t x) { for (int s = 24; s >= 0; s -= 8) v.push_back((unsigned char)(x >> s)); }
void png_chunk(FILE* f, const char type[4], const std::vector<unsigned char>& data) {
    std::vector<unsigned char> head;
    put_be32(head, (uint32_t)data.size());
    std::fwrite(head.data(), 1, 4, f);
    std::fwrite(type, 1, 4, f);
    if (!data.empty()) std::fwrite(data.data(), 1, data.size(), f);
    uint32_t crc = crc32_update(0xffffffffu, (const unsigned char*)type, 4);
    crc = crc32_update(crc, data.data(), data.size()) ^ 0xffffffffu;
    std::vector<unsigned char> tail;
    put_be32(tail, crc);
    std::fwrite(tail.data(), 1, 4, f);
}
uint64_t fnv1a64(const void* data, size_t n) {
    const unsigned char* p = (const unsigned char*)data;
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 0x100000001b3ull; }
    return h;
}
}  // namespace

// The frame buffer is the reference's (Camera.cpp:79, Camera.h:35): W*H little-endian 0x00RRGGBB words, row 0 at the
// BOTTOM (what StretchDIBits consumes, WinMain.cpp:217); image files are written top row first.
// .ppm -> binary P6; .png -> 8-bit RGB, zlib "stored" blocks (no compression library needed, any viewer reads it).
std::string save_frame(const char* path, const uint32_t* bgra, int W, int H) {
    const std::string name(path);
    const bool png = name.size() >= 4 && name.compare(name.size() - 4, 4, ".png") == 0;
    FILE* f = std::fopen(path, "wb");
    if (!f) return std::string("write_frame: cannot open ") + path;
    std::vector<unsigned char> row((size_t)W * 3);
    auto fill_row = [&](int y) {
        const uint32_t* src = bgra + (size_t)(H - 1 - y) * W;
        for (int x = 0; x < W; x++) { row[3 * x] = (unsigned char)(src[x] >> 16); row[3 * x + 1] = (unsigned char)(src[x] >> 8); row[3 * x + 2] = (unsigned char)src[x]; }
    };
    if (!png) {
        std::fprintf(f, "P6\n%d %d\n255\n", W, H);
        for (int y = 0; y < H; y++) { fill_row(y); std::fwrite(row.data(), 1, row.size(), f); }
        std::fclose(f);
        return "";
    }
    const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::fwrite(sig, 1, 8, f);
    std::vector<unsigned char> ihdr;
    put_be32(ihdr, (uint32_t)W); put_be32(ihdr, (uint32_t)H);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);  // 8-bit RGB
    png_chunk(f, "IHDR", ihdr);
    std::vector<unsigned char> raw;  // filter byte 0 + RGB per row
    raw.reserve((size_t)H * (row.size() + 1));
    for (int y = 0; y < H; y++) { fill_row(y); raw.push_back(0); raw.insert(raw.end(), row.begin(), row.end()); }
    std::vector<unsigned char> z;
    z.push_back(0x78); z.push_back(0x01);
    uint32_t a = 1, b = 0;  // Adler-32 of the raw stream
    for (size_t at = 0; at < raw.size() || at == 0;) {
        const size_t n = std::min<size_t>(65535, raw.size() - at);
        const bool last = at + n >= raw.size();
        z.push_back(last ? 1 : 0);
        z.push_back((unsigned char)(n & 0xff)); z.push_back((unsigned char)(n >> 8));
        z.push_back((unsigned char)(~n & 0xff)); z.push_back((unsigned char)((~n >> 8) & 0xff));
        z.insert(z.end(), raw.begin() + (long)at, raw.begin() + (long)(at + n));
        for (size_t i = at; i < at + n; i++) { a = (a + raw[i]) % 65521u; b = (b + a) % 65521u; }
        at += n;
        if (last) break;
    }
    put_be32(z, (b << 16) | a);
    png_chunk(f, "IDAT", z);
    png_chunk(f, "IEND", std::vector<unsigned char>());
    std::fclose(f);
    return "";
}

// Tree cache: the built tree of one mesh as a flat little-endian blob, tied to the mesh by a hash of its points and
// protected by a hash of its own payload.
//   "RTBKD2\0\0" | u64 num_tri | u64 num_nodes | u64 fnv1a64(points9) | u64 fnv1a64 chained over the six arrays |
//   bounds[6N] f32 | left[N] i32 | tri[N] i32 | s1[N] f32 | s2[N] f32 | cut_flag[N] u8
// ("RTBKD1" files, which lack the payload hash, are still read; both go through the same structural validation.)
namespace {
uint64_t tree_payload_hash(const HostTree& T) {
    const size_t N = (size_t)T.num_nodes;
    uint64_t h = fnv1a64(T.bounds.data(), 4 * 6 * N);
    const struct { const void* p; size_t n; } parts[5] = {{T.left.data(), 4 * N}, {T.tri.data(), 4 * N}, {T.s1.data(), 4 * N}, {T.s2.data(), 4 * N},
                                                           {T.cut_flag.data(), N}};
    for (const auto& part : parts) {  // chain: hash of (previous hash, next array)
        const uint64_t g = fnv1a64(part.p, part.n);
        const uint64_t pair[2] = {h, g};
        h = fnv1a64(pair, sizeof pair);
    }
    return h;
}
// Everything the render kernels rely on: children lie behind their parent (so one forward pass sees every node after
// its parent), every node but the root has exactly ONE parent, leaves name a triangle of the mesh and every triangle is
// named by exactly one leaf, and no node lies deeper than the traversal stack can hold (kMaxTreeDepth: the kernels'
// stack has kStackDepth = 40 entries and a descent pushes at most one entry per level).
constexpr int kMaxTreeDepth = 38;
bool tree_structure_ok(const HostTree& T, int64_t num_tri) {
    const size_t N = (size_t)T.num_nodes;
    std::vector<uint8_t> depth(N, 0), parents(N, 0), tri_seen((size_t)num_tri, 0);
    for (size_t i = 0; i < N; i++) {
        if (i > 0 && parents[i] != 1) return false;  // unreachable, or reachable along two paths
        const int32_t l = T.left[i];
        if (l < 0) {
            const int32_t t = T.tri[i];
            if (t < 0 || t >= num_tri || tri_seen[(size_t)t]++) return false;
            continue;
        }
        if ((size_t)l + 1 >= N || (size_t)l <= i || T.cut_flag[i] >= 6 || depth[i] >= kMaxTreeDepth) return false;
        for (size_t c = (size_t)l; c <= (size_t)l + 1; c++) {
            if (parents[c]++) return false;
            depth[c] = (uint8_t)(depth[i] + 1);
        }
    }
    return true;
}
}  // namespace
std::string save_tree(const char* path, const HostTree& T, const float* points9) {
    FILE* f = std::fopen(path, "wb");
    if (!f) return std::string("save_tree: cannot open ") + path;
    const uint64_t head[4] = {(uint64_t)T.num_tri, (uint64_t)T.num_nodes, fnv1a64(points9, sizeof(float) * 9 * (size_t)T.num_tri), tree_payload_hash(T)};
    const size_t N = (size_t)T.num_nodes;
    bool ok = std::fwrite("RTBKD2\0\0", 1, 8, f) == 8 && std::fwrite(head, 8, 4, f) == 4;
    ok = ok && std::fwrite(T.bounds.data(), 4, 6 * N, f) == 6 * N && std::fwrite(T.left.data(), 4, N, f) == N;
    ok = ok && std::fwrite(T.tri.data(), 4, N, f) == N && std::fwrite(T.s1.data(), 4, N, f) == N && std::fwrite(T.s2.data(), 4, N, f) == N;
    ok = ok && std::fwrite(T.cut_flag.data(), 1, N, f) == N;
    std::fclose(f);
    return ok ? "" : std::string("save_tree: short write to ") + path;
}
std::string load_tree(const char* path, const float* points9, int64_t num_tri, HostTree& T) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return std::string("load_tree: cannot open ") + path;
    char magic[8];
    uint64_t head[4] = {0, 0, 0, 0};
    std::string err;
    int version = 0;
    if (std::fread(magic, 1, 8, f) == 8) version = std::memcmp(magic, "RTBKD2\0\0", 8) == 0 ? 2 : std::memcmp(magic, "RTBKD1\0\0", 8) == 0 ? 1 : 0;
    if (version == 0 || std::fread(head, 8, (size_t)(version == 2 ? 4 : 3), f) != (size_t)(version == 2 ? 4 : 3)) err = "load_tree: not a tree file";
    else if ((int64_t)head[0] != num_tri || head[1] != 2 * head[0] - 1) err = "load_tree: the file holds a tree of a different triangle count";
    else if (head[2] != fnv1a64(points9, sizeof(float) * 9 * (size_t)num_tri)) err = "load_tree: the file belongs to a different mesh (point hash mismatch)";
    if (err.empty()) {
        const size_t N = (size_t)head[1];
        T.num_tri = num_tri; T.num_nodes = (int64_t)N;
        T.bounds.resize(6 * N); T.left.resize(N); T.tri.resize(N); T.s1.resize(N); T.s2.resize(N); T.cut_flag.resize(N);
        bool ok = std::fread(T.bounds.data(), 4, 6 * N, f) == 6 * N && std::fread(T.left.data(), 4, N, f) == N;
        ok = ok && std::fread(T.tri.data(), 4, N, f) == N && std::fread(T.s1.data(), 4, N, f) == N && std::fread(T.s2.data(), 4, N, f) == N;
        ok = ok && std::fread(T.cut_flag.data(), 1, N, f) == N;
        if (!ok) err = "load_tree: truncated file";
        else if (version == 2 && head[3] != tree_payload_hash(T)) err = "load_tree: damaged file (payload hash mismatch)";
        else if (!tree_structure_ok(T, num_tri)) err = "load_tree: damaged file (not a tree over this mesh, or deeper than the traversal stack)";
        T.seconds_sort = T.seconds_partition = 0.0;
    }
    std::fclose(f);
    return err;
}

// =================================================================================================
// procedural stand-in mesh
// =================================================================================================
namespace {

inline uint32_t hash3(int32_t x, int32_t y, int32_t z, uint32_t seed) {
    uint32_t h = seed * 0x9E3779B1u;
    h ^= (uint32_t)x * 0x85EBCA77u; h = (h << 13) | (h >> 19); h *= 0xC2B2AE3Du;
    h ^= (uint32_t)y * 0x27D4EB2Fu; h = (h << 13) | (h >> 19); h *= 0x165667B1u;
    h ^= (uint32_t)z * 0x9E3779B1u; h = (h << 13) | (h >> 19); h *= 0x85EBCA77u;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return h;
}
// trilinear value noise in [-1,1]; only + - * floor, so it is reproducible everywhere
inline double value_noise(double x, double y, double z, uint32_t seed) {
    double fx = std::floor(x), fy = std::floor(y), fz = std::floor(z);
    int32_t ix = (int32_t)fx, iy = (int32_t)fy, iz = (int32_t)fz;
    double tx = x - fx, ty = y - fy, tz = z - fz;
    tx = tx * tx * (3.0 - 2.0 * tx); ty = ty * ty * (3.0 - 2.0 * ty); tz = tz * tz * (3.0 - 2.0 * tz);
    double acc = 0.0;
    for (int c = 0; c < 8; c++) {
        int dx = c & 1, dy = (c >> 1) & 1, dz = (c >> 2) & 1;
        double wgt = (dx ? tx : 1.0 - tx) * (dy ? ty : 1.0 - ty) * (dz ? tz : 1.0 - tz);
        double val = (double)(hash3(ix + dx, iy + dy, iz + dz, seed) >> 8) * (1.0 / 8388607.5) - 1.0;
        acc += wgt * val;
    }
    return acc;
}

}  // namespace

void make_geodesic(int nu, float radius, const float center[3], float displacement, uint32_t seed,
                   std::vector<float>& points9) {
    const double t = (1.0 + std::sqrt(5.0)) / 2.0;
    const double ico[12][3] = {{-1, t, 0}, {1, t, 0}, {-1, -t, 0}, {1, -t, 0}, {0, -1, t}, {0, 1, t},
                               {0, -1, -t}, {0, 1, -t}, {t, 0, -1}, {t, 0, 1}, {-t, 0, -1}, {-t, 0, 1}};
    const int faces[20][3] = {{0, 11, 5}, {0, 5, 1}, {0, 1, 7}, {0, 7, 10}, {0, 10, 11}, {1, 5, 9}, {5, 11, 4},
                              {11, 10, 2}, {10, 7, 6}, {7, 1, 8}, {3, 9, 4}, {3, 4, 2}, {3, 2, 6}, {3, 6, 8},
                              {3, 8, 9}, {4, 9, 5}, {2, 4, 11}, {6, 2, 10}, {8, 6, 7}, {9, 8, 1}};
    points9.assign((size_t)20 * nu * nu * 9, 0.0f);
    auto vertex = [&](const int* f, int i, int j, float* out) {
        const double a = (double)(nu - i - j) / nu, b = (double)i / nu, c = (double)j / nu;
        double d[3];
        for (int k = 0; k < 3; k++) d[k] = a * ico[f[0]][k] + b * ico[f[1]][k] + c * ico[f[2]][k];
        const double inv = 1.0 / std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        for (int k = 0; k < 3; k++) d[k] *= inv;
        // snap the direction so that a vertex shared by two faces is displaced identically
        for (int k = 0; k < 3; k++) d[k] = std::floor(d[k] * 1048576.0 + 0.5) / 1048576.0;
        const double nz = 0.65 * value_noise(d[0] * 3.0 + 7.3, d[1] * 3.0 + 1.9, d[2] * 3.0 + 4.1, seed) +
                          0.35 * value_noise(d[0] * 11.0 + 0.5, d[1] * 11.0 + 8.2, d[2] * 11.0 + 2.7, seed ^ 0x5bd1e995u);
        const double r = (double)radius * (1.0 + (double)displacement * nz);
        for (int k = 0; k < 3; k++) out[k] = (float)((double)center[k] + r * d[k]);
    };
    size_t tri = 0;
    for (int f = 0; f < 20; f++) {
        for (int i = 0; i < nu; i++) {
            for (int j = 0; i + j < nu; j++) {
                float* o = &points9[9 * tri++];
                vertex(faces[f], i, j, o); vertex(faces[f], i + 1, j, o + 3); vertex(faces[f], i, j + 1, o + 6);
                if (i + j < nu - 1) {
                    float* q = &points9[9 * tri++];
                    vertex(faces[f], i + 1, j, q); vertex(faces[f], i + 1, j + 1, q + 3); vertex(faces[f], i, j + 1, q + 6);
                }
            }
        }
    }
}

// =================================================================================================
// tree build
// =================================================================================================
namespace {

inline float min3(float a, float b, float c) { float m = b < c ? b : c; return a < m ? a : m; }
inline float max3(float a, float b, float c) { float m = b > c ? b : c; return a > m ? a : m; }

template <class F>
void parallel_for(int threads, int64_t count, int64_t grain, F&& body) {
    if (threads <= 1 || count <= grain) {
        for (int64_t i = 0; i < count; i++) body(i);
        return;
    }
    std::atomic<int64_t> next{0};
    auto worker = [&]() {
        for (;;) {
            int64_t b = next.fetch_add(grain);
            if (b >= count) return;
            int64_t e = std::min(count, b + grain);
            for (int64_t i = b; i < e; i++) body(i);
        }
    };
    std::vector<std::thread> pool;
    int nt = (int)std::min<int64_t>(threads, (count + grain - 1) / grain);
    for (int t = 1; t < nt; t++) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
}

// list numbering of the reference (Trixel.h:217-236): 0=x1 1=y1 2=z1 3=x0 4=y0 5=z0
constexpr int kScanOrder[6] = {0, 3, 1, 4, 2, 5};  // Trixel.h:172-193

}  // namespace

void build_tree(const float* points9, int64_t n, HostTree& T, int threads) {
    using clock = std::chrono::steady_clock;
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    const auto t_begin = clock::now();
    const int64_t N = 2 * n - 1;
    T.num_tri = n; T.num_nodes = N;
    T.bounds.assign((size_t)N * 6, 0.0f);
    T.left.assign((size_t)N, -1);
    T.tri.assign((size_t)N, -1);
    T.cut_flag.assign((size_t)N, 0);
    T.s1.assign((size_t)N, 0.0f);
    T.s2.assign((size_t)N, 0.0f);

    // per-triangle AABB keys (read_ply.cpp:127-134)
    std::vector<float> key[6];
    std::vector<int32_t> order[6], rank[6], scratch[6];
    for (int k = 0; k < 6; k++) { key[k].resize((size_t)n); order[k].resize((size_t)n); rank[k].resize((size_t)n); scratch[k].resize((size_t)n); }
    parallel_for(threads, n, 1 << 16, [&](int64_t i) {
        const float* p = points9 + 9 * i;
        key[0][i] = max3(p[0], p[3], p[6]); key[3][i] = min3(p[0], p[3], p[6]);
        key[1][i] = max3(p[1], p[4], p[7]); key[4][i] = min3(p[1], p[4], p[7]);
        key[2][i] = max3(p[2], p[5], p[8]); key[5][i] = min3(p[2], p[5], p[8]);
    });
    // six sorted lists.  The reference's top-down merge takes the right run on ties (sort.h:31-54),
    // i.e. equal keys end up ordered by DESCENDING original index.
    parallel_for(std::min(threads, 6), 6, 1, [&](int64_t k) {
        auto& ord = order[k];
        const float* ky = key[k].data();
        for (int64_t i = 0; i < n; i++) ord[(size_t)i] = (int32_t)i;
        std::sort(ord.begin(), ord.end(), [ky](int32_t a, int32_t b) {
            const float ka = ky[a], kb = ky[b];
            return ka < kb || (!(kb < ka) && a > b);
        });
        auto& rk = rank[k];
        for (int64_t i = 0; i < n; i++) rk[(size_t)ord[(size_t)i]] = (int32_t)i;
    });
    const auto t_sorted = clock::now();

    // level-synchronous partition.  BFS numbering: the children of the interior nodes of one level,
    // taken in node order, are the consecutive nodes of the next level (Trixel.h:143,329-352).
    std::vector<int64_t> lo((size_t)N), hi((size_t)N);
    std::vector<int32_t> parent((size_t)N, 0);
    lo[0] = 0; hi[0] = n - 1;
    auto set_bounds = [&](int64_t node, int64_t a, int64_t b) {
        float* B = &T.bounds[(size_t)node * 6];
        B[0] = key[3][order[3][a]]; B[1] = key[0][order[0][b]];
        B[2] = key[4][order[4][a]]; B[3] = key[1][order[1][b]];
        B[4] = key[5][order[5][a]]; B[5] = key[2][order[2][b]];
    };
    set_bounds(0, 0, n - 1);
    T.cut_flag[0] = 5;  // Trixel.h:152 (only visible if the root is itself a leaf)
    int64_t level_begin = 0, level_end = 1;
    std::vector<int64_t> child_base;
    while (level_begin < level_end) {
        const int64_t count = level_end - level_begin;
        child_base.assign((size_t)count + 1, 0);
        // choose the split list of every node of this level
        parallel_for(threads, count, 4096, [&](int64_t q) {
            const int64_t node = level_begin + q, l = lo[node], r = hi[node];
            if (r == l) {
                T.cut_flag[node] = T.cut_flag[parent[node]];  // Trixel.h:194
                T.tri[node] = order[0][l];                    // Trixel.h:202
                child_base[(size_t)q + 1] = 0;
                return;
            }
            float best = key[0][order[0][r]] - key[0][order[0][l]];
            int cut = 0;
            for (int s = 1; s < 6; s++) {
                const int k = kScanOrder[s];
                const float spread = key[k][order[k][r]] - key[k][order[k][l]];
                if (spread > best) { best = spread; cut = k; }
            }
            T.cut_flag[node] = (uint8_t)cut;
            child_base[(size_t)q + 1] = 2;
        });
        for (int64_t q = 0; q < count; q++) child_base[(size_t)q + 1] += child_base[(size_t)q];
        const int64_t next_begin = level_end, next_end = level_end + child_base[(size_t)count];
        // stable partition of the five other lists of every interior node (Trixel.h:214-327);
        // task = (node, list), ranges of different nodes are disjoint
        parallel_for(threads, count * 6, 64, [&](int64_t task) {
            const int64_t q = task / 6, node = level_begin + q;
            const int k = (int)(task % 6);
            const int64_t l = lo[node], r = hi[node];
            if (r == l) return;
            const int cut = T.cut_flag[node];
            if (k == cut) return;
            const int64_t m = l + (r - l) / 2;
            const int32_t* cut_rank = rank[cut].data();
            int32_t* ord = order[k].data();
            int32_t* tmp = scratch[k].data();
            int64_t a = l, b = m + 1;
            for (int64_t i = l; i <= r; i++) {
                const int32_t t = ord[i];
                if (cut_rank[t] <= m) tmp[a++] = t; else tmp[b++] = t;
            }
            int32_t* rk = rank[k].data();
            for (int64_t i = l; i <= r; i++) { ord[i] = tmp[i]; rk[tmp[i]] = (int32_t)i; }
        });
        // create the children
        parallel_for(threads, count, 4096, [&](int64_t q) {
            const int64_t node = level_begin + q, l = lo[node], r = hi[node];
            if (r == l) return;
            const int64_t m = l + (r - l) / 2;
            const int64_t cl = next_begin + child_base[(size_t)q], cr = cl + 1;
            T.left[node] = (int32_t)cl;
            lo[cl] = l; hi[cl] = m; parent[cl] = (int32_t)node;
            lo[cr] = m + 1; hi[cr] = r; parent[cr] = (int32_t)node;
            set_bounds(cl, l, m);
            set_bounds(cr, m + 1, r);
            const int axis = T.cut_flag[node] % 3;  // Trixel.h:353-376
            T.s1[node] = T.bounds[(size_t)cl * 6 + 2 * axis + 1];
            T.s2[node] = T.bounds[(size_t)cr * 6 + 2 * axis];
        });
        level_begin = next_begin;
        level_end = next_end;
    }
    const auto t_done = clock::now();
    T.seconds_sort = std::chrono::duration<double>(t_sorted - t_begin).count();
    T.seconds_partition = std::chrono::duration<double>(t_done - t_sorted).count();
}

// =================================================================================================
// camera basis
// =================================================================================================
namespace {

// vector.cpp:13-26 (host rsqrt: magic constant + 8 Newton steps, 32-bit integer view)
inline float host_rsqrt(float s) {
    const float half = 0.5f * s;
    int32_t bits;
    std::memcpy(&bits, &half, 4);
    bits = 0x5f375a86 - (bits >> 1);
    float g;
    std::memcpy(&g, &bits, 4);
    for (int it = 0; it < 8; it++) g = g * (1.5f - half * g * g);
    return g;
}
// Vector.h:116-124
inline void normalize(Vec4& v) {
    float s = v.x * v.x + v.y * v.y + v.z * v.z;
    s = host_rsqrt(s);
    v.x *= s; v.y *= s; v.z *= s;
    v.w = 1 / s;
}
// vector.cpp:31-36
inline Vec4 cross(const Vec4& a, const Vec4& b) {
    Vec4 r = a;
    r.x = a.y * b.z - a.z * b.y;
    r.y = a.z * b.x - a.x * b.z;
    r.z = a.x * b.y - a.y * b.x;
    return r;
}

}  // namespace

void camera_basis(int32_t W, int32_t H, float f_w, float f_h, float fclen, const float pos[3], const float la[3],
                  const float up[3], CameraBasis& c) {
    c.W = W; c.H = H;
    std::memcpy(c.pos, pos, 12);
    const float pix_w = f_w / (float)W, pix_h = f_h / (float)H;  // Camera.cpp:16-17
    Vec4 n{la[0] - pos[0], la[1] - pos[1], la[2] - pos[2], 1.0f};
    normalize(n);
    Vec4 upv{up[0], up[1], up[2], 1.0f};
    normalize(upv);
    Vec4 v = cross(n, cross(upv, n));  // Camera.cpp:38-39: up x n, then n x (up x n)
    normalize(v);
    Vec4 u = cross(v, n);  // Camera.cpp:52
    c.n[0] = n.x; c.n[1] = n.y; c.n[2] = n.z;
    c.v[0] = v.x; c.v[1] = v.y; c.v[2] = v.z;
    c.u[0] = u.x; c.u[1] = u.y; c.u[2] = u.z;
    float ay = (float)((uint32_t)H >> 1), ax = (float)((uint32_t)W >> 1);  // Camera.cpp:61-63
    if (!(H & 1)) ay -= 0.5f;
    if (!(W & 1)) ax -= 0.5f;
    for (int k = 0; k < 3; k++) {
        c.v_mod[k] = c.v[k] * pix_h;
        c.u_mod[k] = c.u[k] * pix_w;
        c.n_mod[k] = (c.n[k] * fclen) - (c.v_mod[k] * ay) - (c.u_mod[k] * ax);  // Camera.cpp:65-67
    }
}

// =================================================================================================
// transform recurrence
// =================================================================================================
void Transform::reset(const float cam_pos[3]) {
    quat = Vec4{0, 0, 0, 1};  // Quaternion.cpp:10
    row[0] = Vec4{1, 0, 0, 0}; row[1] = Vec4{0, 1, 0, 0}; row[2] = Vec4{0, 0, 1, 0};
    init_face = Vec4{-cam_pos[0], -cam_pos[1], -cam_pos[2], 1.0f};  // Camera.cpp:131-132
    cur_face = init_face;
}

namespace {

// quaternion accumulate + matrix, vector.cpp:40-58
inline void accumulate(Transform& T, const Vec4& s) {
    const float a = T.quat.x, b = T.quat.y, c = T.quat.z, d = T.quat.w;
    T.quat.x = b * s.z - c * s.y + a * s.w + d * s.x;
    T.quat.y = c * s.x - a * s.z + b * s.w + d * s.y;
    T.quat.z = a * s.y - b * s.x + c * s.w + d * s.z;
    T.quat.w = d * s.w - a * s.x - b * s.y - c * s.z;
    const float i = T.quat.x, j = T.quat.y, k = T.quat.z, w = T.quat.w;
    T.row[0].x = 1 - 2 * j * j - 2 * k * k; T.row[0].y = 2 * i * j - 2 * k * w; T.row[0].z = 2 * i * k + 2 * j * w;
    T.row[1].x = 2 * i * j + 2 * k * w; T.row[1].y = 1 - 2 * i * i - 2 * k * k; T.row[1].z = 2 * j * k - 2 * i * w;
    T.row[2].x = 2 * i * k - 2 * j * w; T.row[2].y = 2 * j * k + 2 * i * w; T.row[2].z = 1 - 2 * i * i - 2 * j * j;
}
// vector.cpp:60-64 with reverse = -1
inline void rotate_reversed(const Transform& T, Vec4& v) {
    const float tx = v.x * -1, ty = v.y * -1, tz = v.z * -1;
    v.x = tx * T.row[0].x + ty * T.row[0].y + tz * T.row[0].z;
    v.y = tx * T.row[1].x + ty * T.row[1].y + tz * T.row[1].z;
    v.z = tx * T.row[2].x + ty * T.row[2].y + tz * T.row[2].z;
}
// Vector.h:89-100: both operands are scaled by their own w first
inline void sub_scaled(Vec4& a, const Vec4& b) {
    a.x = (a.x * a.w) - (b.x * b.w); a.y = (a.y * a.w) - (b.y * b.w); a.z = (a.z * a.w) - (b.z * b.w); a.w = 1.0f;
}
inline void add_scaled(Vec4& a, const Vec4& b) {
    a.x = (a.x * a.w) + (b.x * b.w); a.y = (a.y * a.w) + (b.y * b.w); a.z = (a.z * a.w) + (b.z * b.w); a.w = 1.0f;
}
inline void negate(Vec4& v) { v.x = -v.x; v.y = -v.y; v.z = -v.z; }

}  // namespace

bool Transform::apply(uint8_t select, float x, float y, float z, float w) {
    Vec4 in{x, y, z, w};
    if (select == 30 || select == 31 || select == 32) {  // Camera.cu:257-287
        sub_scaled(init_face, in);
        rotate_reversed(*this, in);
        row[0].w += in.w * in.x;
        row[1].w += in.w * in.y;
        row[2].w += in.w * in.z;
        normalize(init_face);
        cur_face = init_face;
        rotate_reversed(*this, cur_face);
        negate(cur_face);
        return true;
    }
    if (select == 10 || select == 11) {  // Camera.cu:288-329
        Vec4 probe = init_face;
        accumulate(*this, in);
        rotate_reversed(*this, probe);
        add_scaled(probe, cur_face);
        row[0].w -= probe.x * probe.w;
        row[1].w -= probe.y * probe.w;
        row[2].w -= probe.z * probe.w;
        sub_scaled(probe, cur_face);
        cur_face = probe;
        normalize(cur_face);
        negate(cur_face);
        return true;
    }
    return false;
}

void Transform::matrix(float m[12]) const {
    for (int r = 0; r < 3; r++) { m[4 * r] = row[r].x; m[4 * r + 1] = row[r].y; m[4 * r + 2] = row[r].z; m[4 * r + 3] = row[r].w; }
}
void Transform::set_matrix(const float m[12]) {
    for (int r = 0; r < 3; r++) { row[r].x = m[4 * r]; row[r].y = m[4 * r + 1]; row[r].z = m[4 * r + 2]; row[r].w = m[4 * r + 3]; }
}


// Fill `count` 32-bit words with `value` on `threads` host threads.  Streaming (non-temporal) stores where the platform
// has them: the destination is a frame buffer that is written once and read later by somebody else, so there is no
// point in first reading its cache lines in (which is what ordinary stores do) -- about half the memory traffic.
namespace {
inline void fill_range(uint32_t* p, uint32_t* const end, uint32_t value) {
#if defined(__x86_64__)
    while (p < end && ((uintptr_t)p & 15u)) *p++ = value;
    const __m128i v = _mm_set1_epi32((int)value);
    for (; p + 16 <= end; p += 16) {
        _mm_stream_si128((__m128i*)p, v); _mm_stream_si128((__m128i*)(p + 4), v);
        _mm_stream_si128((__m128i*)(p + 8), v); _mm_stream_si128((__m128i*)(p + 12), v);
    }
    _mm_sfence();
#endif
    while (p < end) *p++ = value;
}
}  // namespace
// The fills of a sweep come back to back, a few hundred megabytes each, so their threads are kept: a small pool that
// lives as long as the process (workers sleep on a condition variable between jobs; a job is a range of task indices
// handed out by an atomic counter, the calling thread works too).  Starting a dozen threads per call cost 10-15 % of a
// 256 MB fill.
namespace {
class FillPool {
public:
    void run(int threads, int64_t tasks, const std::function<void(int64_t)>& body) {
        std::lock_guard<std::mutex> one_job_at_a_time(job_lock_);
        const int helpers = (int)std::min<int64_t>(std::max(0, threads - 1), std::max<int64_t>(0, tasks - 1));
        {
            std::unique_lock<std::mutex> lk(m_);
            while ((int)workers_.size() < helpers) workers_.emplace_back([this, id = (int)workers_.size()] { work(id); });
            body_ = &body; tasks_ = tasks; next_.store(0); active_ = helpers; wanted_ = helpers; generation_++;
        }
        wake_.notify_all();
        drain();
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return active_ == 0; });
        body_ = nullptr;
    }
private:
    void drain() {
        for (;;) {
            const int64_t t = next_.fetch_add(1);
            if (t >= tasks_) return;
            (*body_)(t);
        }
    }
    void work(int id) {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_);
                wake_.wait(lk, [&] { return generation_ != seen; });
                seen = generation_;
                if (id >= wanted_) continue;  // this job wants fewer helpers than the pool holds
            }
            drain();
            std::unique_lock<std::mutex> lk(m_);
            if (--active_ == 0) done_.notify_one();
        }
    }
    std::mutex job_lock_, m_;
    std::condition_variable wake_, done_;
    std::vector<std::thread> workers_;  // never joined: they sleep until the process ends
    const std::function<void(int64_t)>* body_ = nullptr;
    int64_t tasks_ = 0;
    std::atomic<int64_t> next_{0};
    int active_ = 0, wanted_ = 0;
    uint64_t generation_ = 0;
};
FillPool& fill_pool() {
    static FillPool* pool = new FillPool();  // intentionally leaked: its sleeping workers must not be torn down by static destructors
    return *pool;
}
}  // namespace
// two buffers in one parallel region (either may be null): the colour and the id frames of a chunk of a sweep
void fill_words2(uint32_t* a, uint32_t value_a, uint32_t* b, uint32_t value_b, size_t count, int threads) {
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    const size_t chunk = (size_t)1 << 18;  // 1 MB per task
    const int64_t per = (int64_t)((count + chunk - 1) / chunk);
    const int64_t tasks = per * ((a ? 1 : 0) + (b ? 1 : 0));
    const std::function<void(int64_t)> body = [&](int64_t task) {
        const bool second = a ? task >= per : true;
        uint32_t* const base = second ? b : a;
        const size_t t = (size_t)(second && a ? task - per : task);
        fill_range(base + t * chunk, base + std::min(count, (t + 1) * chunk), second ? value_b : value_a);
    };
    if (tasks <= 1 || threads <= 1) {
        for (int64_t t = 0; t < tasks; t++) body(t);
        return;
    }
    fill_pool().run(threads, tasks, body);
}
void fill_words(uint32_t* dst, size_t count, uint32_t value, int threads) { fill_words2(dst, value, nullptr, 0, count, threads); }

}  // namespace rtb
