// goicp_b200.hpp -- C++ mirror of the reference's `class GoICP` (src/goicp/jly_goicp.h:82-141)
// on top of the C ABI (include/goicp_b200.h).  Same public member names and call protocol as the
// reference uses in src/main.cpp:47-59,154-159, so existing callers switch by changing the
// include and linking libgoicp_b200.so:
//
//     goicp_b200::GoICP goicp(mse_threshold);
//     goicp.pModel = model.data(); goicp.Nm = model.size();     // packed xyz floats (glm::vec3-compatible)
//     goicp.pData  = data.data();  goicp.Nd = data.size();
//     goicp.BuildDT();
//     goicp.Register();
//     use(goicp.optR, goicp.optT, goicp.optError);
//
// Differences from the reference: Register() runs on the GPU; failures throw std::runtime_error
// (the reference exit()s); the three process globals it needs (goicp_finished, mse_threshold,
// sse_threshold; jly_goicp.cpp:36-38) are replaced by Cancel() and the constructor argument.
#pragma once
#include <cstring>
#include <stdexcept>
#include <string>
#include "goicp_b200.h"

namespace goicp_b200 {

struct POINT3D { float x, y, z; };                     // layout-compatible with glm::vec3 (jly_goicp.h:42)
struct ROTNODE { float a, b, c, w, ub, lb; int l; };   // jly_goicp.h:44-58
struct TRANSNODE { float x, y, z, w, ub, lb; };        // jly_goicp.h:60-72

// just enough of Geiger's `Matrix` (matrix.h) for optR / optT: row-major val[r][c]
template <int M, int N>
struct MatrixMN {
    float val[M][N];
    MatrixMN() { std::memset(val, 0, sizeof val); }
    static MatrixMN eye() { MatrixMN m; for (int i = 0; i < (M < N ? M : N); i++) m.val[i][i] = 1; return m; }
};

class GoICP {
public:
    int Nm = 0, Nd = 0;
    POINT3D* pModel = nullptr;
    POINT3D* pData = nullptr;
    ROTNODE initNodeRot;
    TRANSNODE initNodeTrans;
    // DT3D's public knobs (jly_3ddt.h:100-111); mode: goicp_dt_mode (default: exact EDT of the reference's seed set -- same
    // registrations as GOICP_DT_REFERENCE, the bit-exact propagation, at 1/200 of its build time)
    struct { int SIZE = 300; double expandFactor = 2.0; int mode = GOICP_DT_EXACT_EDT_REFSEED; } dt;
    int numerics = GOICP_NUM_STRICT;                    // goicp_numerics flags
    int searchMode = GOICP_SEARCH_GOICP;                // goicp_search_mode
    float MSEThresh, SSEThresh = 0, optError = 1e+10f;
    float trimFraction = 0.0f;
    bool doTrim = true;
    MatrixMN<3, 3> optR = MatrixMN<3, 3>::eye();
    MatrixMN<3, 1> optT;
    bool finished = false;
    goicp_result result{};                              // counters, certificate, timings

    explicit GoICP(float mse_threshold, int device = 0) : MSEThresh(mse_threshold), device_(device)
    {
        goicp_params p; goicp_default_params(&p);
        initNodeRot = {p.rot_cube[0], p.rot_cube[1], p.rot_cube[2], p.rot_cube[3], 0, 0, 0};
        initNodeTrans = {p.trans_cube[0], p.trans_cube[1], p.trans_cube[2], p.trans_cube[3], 0, 0};
    }
    ~GoICP() { if (h_) goicp_destroy(h_); }
    GoICP(const GoICP&) = delete;
    GoICP& operator=(const GoICP&) = delete;

    void BuildDT() { open(); check(goicp_build_dt(h_)); }                      // jly_goicp.cpp:75-90
    float Register()                                                            // jly_goicp.cpp:569-585
    {
        open();
        check(goicp_register(h_, &result));
        for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) optR.val[i][j] = result.R[3 * i + j]; optT.val[i][0] = result.t[i]; }
        optError = result.sse; SSEThresh = result.sse_thresh; finished = true;
        return optError;
    }
    void Cancel() { if (h_) goicp_cancel(h_); }                                // replaces the global goicp_finished
    goicp_snapshot Poll() { goicp_snapshot s{}; if (h_) goicp_poll(h_, &s); return s; }

private:
    goicp_handle* h_ = nullptr;
    int device_;
    void open()
    {
        if (h_) return;
        if (!pModel || !pData || Nm <= 0 || Nd <= 0) throw std::runtime_error("GoICP: pModel/Nm/pData/Nd not set");
        goicp_params p; goicp_default_params(&p);
        p.mse_threshold = MSEThresh; p.trim_fraction = trimFraction; p.do_trim = doTrim ? 1 : 0;
        p.dt_size = dt.SIZE; p.dt_expand = dt.expandFactor; p.dt_mode = dt.mode; p.device = device_; p.numerics = numerics; p.search_mode = searchMode;
        p.rot_cube[0] = initNodeRot.a; p.rot_cube[1] = initNodeRot.b; p.rot_cube[2] = initNodeRot.c; p.rot_cube[3] = initNodeRot.w;
        p.trans_cube[0] = initNodeTrans.x; p.trans_cube[1] = initNodeTrans.y; p.trans_cube[2] = initNodeTrans.z; p.trans_cube[3] = initNodeTrans.w;
        if (goicp_create(&p, &h_)) throw std::runtime_error("goicp_create failed");
        check(goicp_set_model(h_, &pModel->x, Nm));
        check(goicp_set_data(h_, &pData->x, Nd));
    }
    void check(int rc) { if (rc) throw std::runtime_error(std::string("goicp_b200: ") + goicp_last_error(h_)); }
};

} // namespace goicp_b200
