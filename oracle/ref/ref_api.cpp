// TEST INFRASTRUCTURE ONLY (oracle/_ref build) -- never linked into the product library.
//
// C-ABI shim + CLI over the UNMODIFIED reference CPU Go-ICP
// (/root/reference/src/goicp/{jly_goicp,jly_3ddt,matrix}.cpp, compiled where they lie by
// oracle/Makefile).  Used to (1) pin the C restatement in oracle/goicp_oracle.c,
// (2) generate tests/golden/*.json, (3) serve as bench.py's `--impl reference` arm.
// Private members of the reference classes are reached with the usual
// `#define private public` trick (layout is unaffected).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <cmath>
#include <ctime>
#include <chrono>
#include <iostream>
#include <sstream>
#include <fstream>
#include <vector>
#include <queue>
#include <random>
#include <string>
#include <algorithm>
#include <stdexcept>
#include <limits>
#include <cassert>

#define private public
#include "goicp/jly_goicp.h"
#undef private

// process globals the reference expects (jly_goicp.cpp:36-38)
bool goicp_finished = false;
float mse_threshold = 1e-3f;
float sse_threshold = 0.f;
extern long long tNodeCount, rNodeCount;   // jly_goicp.cpp:34-35
extern long long ref_select_calls;         // ref_tu_goicp.cpp

static double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct RefHandle {
    GoICP* g;
    std::vector<POINT3D> model, data;
    bool dt_built;
};

extern "C" {

// ---- deterministic subsample (replaces common.cpp:171-184's random_device) ------------
// keep point i iff dis(gen) <= sub && kept < (size_t)(n*sub); out must hold n*3 floats.
size_t ref_subsample(const float* xyz, size_t n, float sub, float resize, unsigned seed, float* out)
{
    std::mt19937 gen(seed);
    std::uniform_real_distribution<float> dis(0.0, 1.0);
    size_t cap = static_cast<size_t>(n * sub), kept = 0;
    for (size_t i = 0; i < n; ++i) {
        if (dis(gen) <= sub && kept < cap) {
            out[3*kept+0] = resize * xyz[3*i+0];
            out[3*kept+1] = resize * xyz[3*i+1];
            out[3*kept+2] = resize * xyz[3*i+2];
            ++kept;
        }
    }
    return kept;
}

// ---- distance transform ---------------------------------------------------------------
// Builds DT3D exactly as GoICP::BuildDT does (jly_goicp.cpp:75-90) and copies out the
// float distance grid in [z][y][x] order plus meta {xMin,yMin,zMin,scale}.
void* ref_dt_build(const float* model_xyz, int nm, int size, double expand)
{
    DT3D* dt = new DT3D();
    dt->SIZE = size;
    dt->expandFactor = expand;
    std::vector<double> x(nm), y(nm), z(nm);
    for (int i = 0; i < nm; i++) { x[i] = model_xyz[3*i]; y[i] = model_xyz[3*i+1]; z[i] = model_xyz[3*i+2]; }
    dt->Build(x.data(), y.data(), z.data(), nm);
    return dt;
}
void ref_dt_meta(void* h, double* meta4) { DT3D* dt = (DT3D*)h; meta4[0]=dt->xMin; meta4[1]=dt->yMin; meta4[2]=dt->zMin; meta4[3]=dt->scale; }
void ref_dt_grid(void* h, float* out)
{
    DT3D* dt = (DT3D*)h; int S = dt->SIZE; size_t k = 0;
    for (int z = 0; z < S; z++) for (int y = 0; y < S; y++) for (int x = 0; x < S; x++) out[k++] = dt->A.data[z][y][x].distance;
}
void ref_dt_vectors(void* h, short* out)   // (v,h,d) per voxel, [z][y][x] order
{
    DT3D* dt = (DT3D*)h; int S = dt->SIZE; size_t k = 0;
    for (int z = 0; z < S; z++) for (int y = 0; y < S; y++) for (int x = 0; x < S; x++) {
        out[k++] = dt->A.data[z][y][x].v; out[k++] = dt->A.data[z][y][x].h; out[k++] = dt->A.data[z][y][x].d; }
}
void ref_dt_distance(void* h, const float* q_xyz, int n, float* out)
{
    DT3D* dt = (DT3D*)h;
    for (int i = 0; i < n; i++) out[i] = dt->Distance(q_xyz[3*i], q_xyz[3*i+1], q_xyz[3*i+2]);
}
void ref_dt_free(void* h) { delete (DT3D*)h; }

// ---- kd-tree NN + ICP (jly_icp3d.hpp) ------------------------------------------------
void* ref_icp_build(const float* model_xyz, int nm)
{
    ICP3D<float>* icp = new ICP3D<float>();
    std::vector<float> m(model_xyz, model_xyz + 3*(size_t)nm);
    icp->Build(m.data(), nm);
    return icp;
}
void ref_icp_nn(void* h, const float* q_xyz, int n, int* idx, float* d2)
{
    ICP3D<float>* icp = (ICP3D<float>*)h;
    for (int i = 0; i < n; i++) {
        size_t ri; float rd;
        icp->kdtree->knnSearch(q_xyz + 3*(size_t)i, 1, &ri, &rd);
        idx[i] = (int)ri; d2[i] = rd;
    }
}
// R (row-major 3x3) and t are in/out; returns the reference's return value.
float ref_icp_run(void* h, const float* data_xyz, int nd, float* R9, float* t3, int max_iter, float err_diff, float trim, int do_trim)
{
    ICP3D<float>* icp = (ICP3D<float>*)h;
    icp->trim_fraction = trim; icp->do_trim = do_trim != 0;
    Matrix R(3,3), t(3,1);
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) R.val[i][j] = R9[3*i+j]; t.val[i][0] = t3[i]; }
    std::vector<float> d(data_xyz, data_xyz + 3*(size_t)nd);
    float e = icp->Run(d.data(), nd, R, t, (size_t)max_iter, err_diff);
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) R9[3*i+j] = R.val[i][j]; t3[i] = t.val[i][0]; }
    return e;
}
void ref_svd3(const float* H9, float* U9, float* W3, float* V9)
{
    Matrix H(3,3,H9), U, W, V;
    H.svd(U, W, V);
    for (int i = 0; i < 3; i++) { W3[i] = W.val[i][0]; for (int j = 0; j < 3; j++) { U9[3*i+j] = U.val[i][j]; V9[3*i+j] = V.val[i][j]; } }
}

// ---- full Go-ICP ----------------------------------------------------------------------
void* ref_goicp_create(const float* model_xyz, int nm, const float* data_xyz, int nd,
                       float mse, float trim, int dt_size, double dt_expand,
                       const float* trans_cube4 /* x,y,z,w or NULL */)
{
    RefHandle* h = new RefHandle();
    h->model.resize(nm); h->data.resize(nd);
    for (int i = 0; i < nm; i++) h->model[i] = POINT3D(model_xyz[3*i], model_xyz[3*i+1], model_xyz[3*i+2]);
    for (int i = 0; i < nd; i++) h->data[i] = POINT3D(data_xyz[3*i], data_xyz[3*i+1], data_xyz[3*i+2]);
    mse_threshold = mse;
    h->g = new GoICP(mse);
    h->g->pModel = h->model.data(); h->g->Nm = nm;
    h->g->pData = h->data.data();   h->g->Nd = nd;
    h->g->trimFraction = trim;
    h->g->dt.SIZE = dt_size; h->g->dt.expandFactor = dt_expand;
    if (trans_cube4) { h->g->initNodeTrans.x = trans_cube4[0]; h->g->initNodeTrans.y = trans_cube4[1];
                       h->g->initNodeTrans.z = trans_cube4[2]; h->g->initNodeTrans.w = trans_cube4[3]; }
    h->dt_built = false;
    return h;
}
void ref_goicp_set_do_trim(void* hh, int v) { ((RefHandle*)hh)->g->doTrim = v != 0; }      // public field of the reference class
double ref_goicp_build_dt(void* hh)
{
    RefHandle* h = (RefHandle*)hh; double t0 = now_s(); h->g->BuildDT(); h->dt_built = true; return now_s() - t0;
}
void* ref_goicp_dt(void* hh) { return &((RefHandle*)hh)->g->dt; }

// out[0..8]=R row-major, out[9..11]=t, out[12]=optError, out[13]=SSEThresh, out[14]=seconds,
// counters[0]=rot pops, [1]=trans pops, [2]=bound evals, [3]=icp calls (not counted: 0), [4]=select calls
float ref_goicp_register(void* hh, double* out, long long* counters)
{
    RefHandle* h = (RefHandle*)hh;
    long long t0n = tNodeCount, r0n = rNodeCount, s0 = ref_select_calls;
    double t0 = now_s();
    float e = h->g->Register();
    double dt = now_s() - t0;
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) out[3*i+j] = h->g->optR.val[i][j]; out[9+i] = h->g->optT.val[i][0]; }
    out[12] = h->g->optError; out[13] = h->g->SSEThresh; out[14] = dt;
    counters[0] = rNodeCount - r0n; counters[1] = tNodeCount - t0n;
    counters[4] = ref_select_calls - s0; counters[2] = -1; counters[3] = -1;
    return e;
}

// One InnerBnB call on a prepared state (after Initialize()): rotate the data by R9, set
// optError, run the ub pass (level<0) or the lb pass with maxRotDis[level].
// out: [0]=returned value, [1..4]= best translation node x,y,z,w (ub pass only)
void ref_goicp_initialize(void* hh) { RefHandle* h = (RefHandle*)hh; h->g->Initialize(); }
float ref_goicp_inner(void* hh, const float* R9, int level, float opt_error, float* out5, long long* counters)
{
    RefHandle* h = (RefHandle*)hh; GoICP* g = h->g;
    for (int i = 0; i < g->Nd; i++) {
        POINT3D& p = g->pData[i];
        g->pDataTemp[i].x = R9[0]*p.x + R9[1]*p.y + R9[2]*p.z;
        g->pDataTemp[i].y = R9[3]*p.x + R9[4]*p.y + R9[5]*p.z;
        g->pDataTemp[i].z = R9[6]*p.x + R9[7]*p.y + R9[8]*p.z;
    }
    g->optError = opt_error;
    long long t0n = tNodeCount, s0 = ref_select_calls;
    TRANSNODE tn; tn.x = tn.y = tn.z = tn.w = 0; tn.ub = tn.lb = 0;
    float v = (level < 0) ? g->InnerBnB(NULL, &tn) : g->InnerBnB(g->maxRotDis[level], NULL);
    out5[0] = v; out5[1] = tn.x; out5[2] = tn.y; out5[3] = tn.z; out5[4] = tn.w;
    counters[0] = tNodeCount - t0n; counters[1] = ref_select_calls - s0;
    return v;
}
// rotation-uncertainty table row (jly_goicp.cpp:148-160), n = Nd floats
void ref_goicp_maxrotdis(void* hh, int level, float* out) { RefHandle* h = (RefHandle*)hh; memcpy(out, h->g->maxRotDis[level], sizeof(float)*h->g->Nd); }
// GoICP::ICP (ICP3D::Run + DT rescoring, jly_goicp.cpp:93-132)
float ref_goicp_icp(void* hh, float* R9, float* t3)
{
    RefHandle* h = (RefHandle*)hh; Matrix R(3,3), t(3,1);
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) R.val[i][j] = R9[3*i+j]; t.val[i][0] = t3[i]; }
    float e = h->g->ICP(R, t);
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) R9[3*i+j] = R.val[i][j]; t3[i] = t.val[i][0]; }
    return e;
}

} // extern "C"

// ---- CLI ------------------------------------------------------------------------------
static std::vector<float> read_f32(const char* path)
{
    FILE* f = fopen(path, "rb"); if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<float> v(sz / 4); if (fread(v.data(), 4, v.size(), f) != v.size()) exit(2); fclose(f); return v;
}
static std::vector<float> read_txt(const char* path)
{
    std::ifstream in(path); int n = 0; in >> n; std::vector<float> v((size_t)n*3);
    for (size_t i = 0; i < v.size(); i++) in >> v[i];
    return v;
}
static uint64_t fnv1a64(const void* p, size_t n)
{
    const unsigned char* b = (const unsigned char*)p; uint64_t h = 1469598103934665603ULL;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ULL; } return h;
}

#ifndef REF_NO_MAIN
int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr,
        "usage: ref_goicp subsample <in.txt> <sub> <resize> <seed> <out.f32>\n"
        "       ref_goicp goicp <model.f32> <data.f32> <mse> [trim=0] [S=300] [tx ty tz tw]\n"
        "       ref_goicp dt <model.f32> <S> <expand> [out.f32]\n"); return 2; }
    std::string cmd = argv[1];
    if (cmd == "subsample") {
        std::vector<float> in = read_txt(argv[2]); std::vector<float> out(in.size());
        size_t k = ref_subsample(in.data(), in.size()/3, atof(argv[3]), atof(argv[4]), (unsigned)atoi(argv[5]), out.data());
        FILE* f = fopen(argv[6], "wb"); fwrite(out.data(), 4, 3*k, f); fclose(f);
        printf("{\"kept\": %zu, \"total\": %zu}\n", k, in.size()/3); return 0;
    }
    if (cmd == "dt") {
        std::vector<float> m = read_f32(argv[2]); int S = atoi(argv[3]);
        double t0 = now_s(); void* dt = ref_dt_build(m.data(), (int)(m.size()/3), S, atof(argv[4])); double el = now_s() - t0;
        std::vector<float> g((size_t)S*S*S); ref_dt_grid(dt, g.data()); double meta[4]; ref_dt_meta(dt, meta);
        float mx = 0; for (float v : g) mx = std::max(mx, v);
        printf("{\"S\": %d, \"xMin\": %.17g, \"yMin\": %.17g, \"zMin\": %.17g, \"scale\": %.17g, \"fnv1a64\": \"%016llx\", \"max\": %.9g, \"build_s\": %.3f}\n",
               S, meta[0], meta[1], meta[2], meta[3], (unsigned long long)fnv1a64(g.data(), g.size()*4), mx, el);
        if (argc > 5) { FILE* f = fopen(argv[5], "wb"); fwrite(g.data(), 4, g.size(), f); fclose(f); }
        return 0;
    }
    if (cmd == "goicp") {
        std::vector<float> m = read_f32(argv[2]), d = read_f32(argv[3]);
        float mse = atof(argv[4]); float trim = argc > 5 ? atof(argv[5]) : 0.f; int S = argc > 6 ? atoi(argv[6]) : 300;
        float tc[4]; bool has_tc = argc > 10; if (has_tc) for (int i = 0; i < 4; i++) tc[i] = atof(argv[7+i]);
        void* h = ref_goicp_create(m.data(), (int)(m.size()/3), d.data(), (int)(d.size()/3), mse, trim, S, 2.0, has_tc ? tc : NULL);
        // the reference narrates on stdout; keep it (tests parse the "Error*:" lines) and print our JSON last
        double dt_s = ref_goicp_build_dt(h);
        double out[15]; long long c[5];
        ref_goicp_register(h, out, c);
        fflush(stdout);
        printf("REFJSON {\"Nm\": %zu, \"Nd\": %zu, \"mse\": %.9g, \"trim\": %.9g, \"S\": %d, \"R\": [", m.size()/3, d.size()/3, mse, trim, S);
        for (int i = 0; i < 9; i++) printf("%.9g%s", out[i], i < 8 ? ", " : "");
        printf("], \"t\": [%.9g, %.9g, %.9g], \"sse\": %.9g, \"sse_thresh\": %.9g, \"register_s\": %.3f, \"dt_build_s\": %.3f, "
               "\"rot_pops\": %lld, \"trans_pops\": %lld, \"select_calls\": %lld}\n",
               out[9], out[10], out[11], out[12], out[13], out[14], dt_s, c[0], c[1], c[4]);
        fflush(stdout); _Exit(0);
    }
    fprintf(stderr, "unknown command %s\n", cmd.c_str()); return 2;
}
#endif
