// host_io.cpp -- config, cloud loaders and the TOML-driven convenience entry point.
//
// Mirrors the reference's src/common.{h,cpp}: `Config` (TOML keys [info].description,
// [io].{target,source,output,visualization}, [params].{mode,trim,subsample,mse_threshold,resize};
// common.cpp:39-74), `load_cloud` for ".txt" (count line + "x y z" rows, common.cpp:148-203) and
// ".ply" (vertex x/y/z as float, ascii or binary_little_endian, extra properties and trailing
// elements ignored; common.cpp:79-146 via tinyply).  Differences, all deliberate: the subsample
// is SEEDED (the reference draws from std::random_device, so its runs cannot be reproduced);
// [io].output is actually written (the reference reads the key and never writes the file);
// failures come back as GOICP_ERR_IO + message instead of C++ exceptions.
// Host logic only; all numerics run on the GPU through the C ABI.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/goicp_b200.h"

namespace {

std::string g_io_error;

std::string trim_ws(const std::string& s)
{
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) a++;
    while (b > a && std::isspace((unsigned char)s[b - 1])) b--;
    return s.substr(a, b - a);
}

// Minimal TOML subset: [table] / [table.sub] headers, key = "string" | number | true/false | [array],
// '#' comments.  Values are kept as strings under "table.key".
std::map<std::string, std::string> parse_toml(const std::string& path)
{
    std::ifstream in(path);
    if (!in) throw std::runtime_error("Unable to open TOML file: " + path);
    std::map<std::string, std::string> kv;
    std::string line, table;
    while (std::getline(in, line)) {
        bool in_str = false; size_t cut = std::string::npos;
        for (size_t i = 0; i < line.size(); i++) { if (line[i] == '"') in_str = !in_str; if (line[i] == '#' && !in_str) { cut = i; break; } }
        if (cut != std::string::npos) line = line.substr(0, cut);
        line = trim_ws(line);
        if (line.empty()) continue;
        if (line.front() == '[' && line.back() == ']' && line.find('=') == std::string::npos) { table = trim_ws(line.substr(1, line.size() - 2)); continue; }
        const size_t eq = line.find('=');
        if (eq == std::string::npos) throw std::runtime_error("TOML syntax error in " + path + ": " + line);
        std::string key = trim_ws(line.substr(0, eq)), val = trim_ws(line.substr(eq + 1));
        if (val.size() >= 2 && val.front() == '"' && val.back() == '"') val = val.substr(1, val.size() - 2);
        kv[table.empty() ? key : table + "." + key] = val;
    }
    return kv;
}

struct Config {                       // common.h:133-176
    std::string description, target, source, output, visualization;
    // defaults of the reference's Config for keys a TOML omits (common.cpp:11-13, 56-60): mode 1, trim false, mse 1e-5.
    // (Config::trim is parsed but never applied to GoICP by the reference -- doTrim stays true, jly_goicp.cpp:62 -- nor here.)
    int mode = 1; bool trim = false; float subsample = 1.0f, mse_threshold = 1e-5f, resize = 1.0f;
};

Config load_config(const std::string& path)
{
    auto kv = parse_toml(path);
    auto need = [&](const char* k) -> std::string { auto it = kv.find(k); if (it == kv.end()) throw std::runtime_error(std::string("missing TOML key: ") + k); return it->second; };
    auto opt = [&](const char* k, const std::string& d) { auto it = kv.find(k); return it == kv.end() ? d : it->second; };
    Config c;
    c.description = opt("info.description", "");
    c.target = need("io.target"); c.source = need("io.source");
    c.output = opt("io.output", ""); c.visualization = opt("io.visualization", "");
    c.mode = std::stoi(opt("params.mode", "1"));
    c.trim = opt("params.trim", "false") == "true";
    c.subsample = std::stof(opt("params.subsample", "1.0"));
    c.mse_threshold = std::stof(opt("params.mse_threshold", "1e-5"));
    c.resize = std::stof(opt("params.resize", "1.0"));
    c.subsample = std::min(1.0f, std::max(0.0f, c.subsample));               // common.cpp:62-64
    c.mse_threshold = std::max(1e-10f, c.mse_threshold);
    return c;
}

// keep point i iff u_i <= subsample and fewer than floor(n*subsample) points were kept so far
// (common.cpp:171-184), with a seeded generator
void subsample_resize(const std::vector<float>& all, float subsample, float resize, unsigned seed, std::vector<float>& out)
{
    const size_t n = all.size() / 3, quota = static_cast<size_t>(n * subsample);
    std::mt19937 gen(seed);
    std::uniform_real_distribution<float> dis(0.0, 1.0);
    out.clear(); out.reserve(quota * 3);
    size_t kept = 0;
    for (size_t i = 0; i < n; i++)
        if (dis(gen) <= subsample && kept < quota) {
            out.push_back(resize * all[3 * i]); out.push_back(resize * all[3 * i + 1]); out.push_back(resize * all[3 * i + 2]);
            kept++;
        }
}

std::vector<float> read_txt(const std::string& path)
{
    std::ifstream in(path);
    if (!in.is_open()) throw std::runtime_error("Unable to open TXT file: " + path);
    long total = 0;
    in >> total;
    if (total <= 0) throw std::runtime_error("Invalid number of points in the TXT file: " + path);
    std::vector<float> v((size_t)total * 3);
    for (size_t i = 0; i < v.size(); i++) if (!(in >> v[i])) throw std::runtime_error("Error reading point data from TXT file: " + path);
    return v;
}

size_t ply_type_size(const std::string& t)
{
    if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
    if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
    if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
    if (t == "double" || t == "float64") return 8;
    throw std::runtime_error("unsupported PLY property type: " + t);
}

std::vector<float> read_ply(const std::string& path)
{
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::runtime_error("Unable to open file: " + path);
    std::string line;
    auto getl = [&]() { if (!std::getline(in, line)) throw std::runtime_error("Error reading PLY file: truncated header"); if (!line.empty() && line.back() == '\r') line.pop_back(); };
    getl();
    if (line != "ply") throw std::runtime_error("Error reading PLY file: not a PLY file: " + path);
    std::string format; long nvert = -1; bool in_vertex = false, vertex_first = true, seen_element = false;
    struct Prop { std::string type, name; size_t off; };
    std::vector<Prop> props; size_t stride = 0;
    for (;;) {
        getl();
        std::istringstream ss(line); std::string tok; ss >> tok;
        if (tok == "format") ss >> format;
        else if (tok == "element") {
            std::string name; long cnt; ss >> name >> cnt;
            in_vertex = name == "vertex";
            if (in_vertex) { nvert = cnt; vertex_first = !seen_element; }
            seen_element = true;
        } else if (tok == "property" && in_vertex) {
            std::string type; ss >> type;
            if (type == "list") throw std::runtime_error("Error reading PLY file: list property on vertex element");
            std::string name; ss >> name;
            props.push_back({type, name, stride}); stride += ply_type_size(type);
        } else if (tok == "end_header") break;
    }
    if (nvert <= 0) throw std::runtime_error("No vertices found in the PLY file.");
    if (!vertex_first) throw std::runtime_error("Error reading PLY file: vertex element is not first");
    int ix = -1, iy = -1, iz = -1;
    for (size_t i = 0; i < props.size(); i++) { if (props[i].name == "x") ix = (int)i; if (props[i].name == "y") iy = (int)i; if (props[i].name == "z") iz = (int)i; }
    if (ix < 0 || iy < 0 || iz < 0) throw std::runtime_error("PLY file missing 'x', 'y', or 'z' vertex properties.");
    std::vector<float> v((size_t)nvert * 3);
    if (format == "ascii") {
        std::vector<double> row(props.size());
        for (long i = 0; i < nvert; i++) {
            for (size_t k = 0; k < props.size(); k++) if (!(in >> row[k])) throw std::runtime_error("Error reading PLY file: truncated vertex data");
            v[3 * i] = (float)row[ix]; v[3 * i + 1] = (float)row[iy]; v[3 * i + 2] = (float)row[iz];
        }
    } else if (format == "binary_little_endian") {
        for (int k : {ix, iy, iz}) if (ply_type_size(props[k].type) != 4 || props[k].type.find("float") == std::string::npos) throw std::runtime_error("Error reading PLY file: x/y/z must be float32");
        std::vector<unsigned char> buf((size_t)nvert * stride);
        in.read((char*)buf.data(), (std::streamsize)buf.size());
        if ((size_t)in.gcount() != buf.size()) throw std::runtime_error("Error reading PLY file: truncated vertex data");
        for (long i = 0; i < nvert; i++) {
            const unsigned char* r = buf.data() + (size_t)i * stride;
            std::memcpy(&v[3 * i], r + props[ix].off, 4); std::memcpy(&v[3 * i + 1], r + props[iy].off, 4); std::memcpy(&v[3 * i + 2], r + props[iz].off, 4);
        }
    } else throw std::runtime_error("Error reading PLY file: unsupported format " + format);
    return v;
}

std::string resolve(const std::string& p, const std::string& toml_path)
{
    if (std::ifstream(p).good()) return p;
    const size_t slash = toml_path.find_last_of('/');
    const std::string alt = (slash == std::string::npos ? std::string(".") : toml_path.substr(0, slash)) + "/" + p;
    return std::ifstream(alt).good() ? alt : p;
}

std::vector<float> load_cloud(const std::string& path, float subsample, float resize, unsigned seed)      // common.cpp:205-228
{
    const size_t dot = path.find_last_of('.');
    if (dot == std::string::npos) throw std::runtime_error("Filepath does not have a valid extension: " + path);
    std::string ext = path.substr(dot + 1);
    std::transform(ext.begin(), ext.end(), ext.begin(), ::tolower);
    std::vector<float> all;
    if (ext == "ply") all = read_ply(path);
    else if (ext == "txt") all = read_txt(path);
    else throw std::runtime_error("Unsupported file extension: " + ext);
    std::vector<float> out;
    subsample_resize(all, subsample, resize, seed, out);
    return out;
}

void write_outputs(const Config& c, const goicp_result& r, const std::vector<float>& model, const std::vector<float>& data)
{
    if (!c.output.empty()) {
        std::ofstream o(c.output);
        if (!o) throw std::runtime_error("Unable to write output file: " + c.output);
        o.precision(9);
        o << "# written by goicp_b200 (x' = R x + t maps source onto target)\n[result]\n";
        o << "R = [[" << r.R[0] << ", " << r.R[1] << ", " << r.R[2] << "], [" << r.R[3] << ", " << r.R[4] << ", " << r.R[5] << "], [" << r.R[6] << ", " << r.R[7] << ", " << r.R[8] << "]]\n";
        o << "t = [" << r.t[0] << ", " << r.t[1] << ", " << r.t[2] << "]\n";
        o << "sse = " << r.sse << "\nmse = " << (data.empty() ? 0.0 : r.sse / (data.size() / 3)) << "\nlower_bound = " << r.best_lb << "\n";
        static const char* paths[] = {"none", "certified", "early_sse_below_thresh", "queue_empty", "cancelled"};
        o << "exit_path = \"" << paths[r.exit_path] << "\"\nrotation_nodes = " << r.rot_pops << "\ntranslation_nodes = " << r.trans_pops << "\nbound_evaluations = " << r.bound_evals << "\nseconds = " << r.seconds_total << "\n";
    }
    if (!c.visualization.empty()) {
        std::ofstream o(c.visualization);
        if (!o) throw std::runtime_error("Unable to write visualization file: " + c.visualization);
        const size_t nm = model.size() / 3, nd = data.size() / 3;
        o << "ply\nformat ascii 1.0\nelement vertex " << nm + nd << "\nproperty float x\nproperty float y\nproperty float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n";
        for (size_t i = 0; i < nm; i++) o << model[3 * i] << " " << model[3 * i + 1] << " " << model[3 * i + 2] << " 90 90 255\n";
        for (size_t i = 0; i < nd; i++) {
            const float* p = &data[3 * i];
            o << r.R[0] * p[0] + r.R[1] * p[1] + r.R[2] * p[2] + r.t[0] << " " << r.R[3] * p[0] + r.R[4] * p[1] + r.R[5] * p[2] + r.t[1] << " "
              << r.R[6] * p[0] + r.R[7] * p[1] + r.R[8] * p[2] + r.t[2] << " 255 120 40\n";
        }
    }
}

} // namespace

extern "C" {

const char* goicp_io_last_error(void) { return g_io_error.c_str(); }

int goicp_load_cloud(const char* path, float subsample, float resize, unsigned seed, float** xyz_out, int* n_out)
{
    try {
        std::vector<float> v = load_cloud(path, subsample, resize, seed);
        float* p = (float*)std::malloc(std::max<size_t>(v.size(), 1) * sizeof(float));
        std::memcpy(p, v.data(), v.size() * sizeof(float));
        *xyz_out = p; *n_out = (int)(v.size() / 3);
        return GOICP_OK;
    } catch (const std::exception& e) { g_io_error = e.what(); return GOICP_ERR_IO; }
}
void goicp_free_cloud(float* xyz) { std::free(xyz); }

int goicp_run_toml(const char* toml_path, unsigned seed_model, unsigned seed_data, goicp_result* out)
{
    if (!toml_path || !out) return GOICP_ERR_INVALID;
    goicp_handle* h = nullptr;
    try {
        const Config c = load_config(toml_path);
        // the same subsample and resize apply to both clouds (main.cpp:34-35)
        const std::vector<float> model = load_cloud(resolve(c.target, toml_path), c.subsample, c.resize, seed_model);
        const std::vector<float> data = load_cloud(resolve(c.source, toml_path), c.subsample, c.resize, seed_data);
        if (model.empty() || data.empty()) throw std::runtime_error("empty point cloud after subsampling");
        goicp_params p; goicp_default_params(&p);
        p.mse_threshold = c.mse_threshold;
        if (c.mode == 4) { p.trans_cube[0] = p.trans_cube[1] = p.trans_cube[2] = -1.0f; p.trans_cube[3] = 2.0f; }   // fgoicp's translation domain [-1,1]^3 (fgoicp.cpp:118-119)
        int rc = goicp_create(&p, &h);
        if (rc) throw std::runtime_error("goicp_create failed");
        auto chk = [&](int code) { if (code) throw std::runtime_error(std::string(goicp_last_error(h))); };
        chk(goicp_set_model(h, model.data(), (int)(model.size() / 3)));
        chk(goicp_set_data(h, data.data(), (int)(data.size() / 3)));
        std::memset(out, 0, sizeof *out);
        if (c.mode >= 3) {                                   // Go-ICP (modes 3 and 4)
            chk(goicp_build_dt(h));
            chk(goicp_register(h, out));
        } else {                                             // ICP only (modes 0-2)
            goicp_icp_result ir; const float I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, z[3] = {0, 0, 0};
            chk(goicp_icp(h, I, z, 0, -1.0f, &ir));
            std::memcpy(out->R, ir.R, sizeof ir.R); std::memcpy(out->t, ir.t, sizeof ir.t); out->sse = ir.err; out->icp_calls = 1;
        }
        write_outputs(c, *out, model, data);
        goicp_destroy(h);
        return GOICP_OK;
    } catch (const std::exception& e) {
        g_io_error = e.what();
        std::fprintf(stderr, "goicp_run_toml: %s\n", e.what());
        if (h) goicp_destroy(h);
        return GOICP_ERR_IO;
    }
}

} // extern "C"
