// src/vFlowB200.cpp  (new file in the reference tree; link with -lfarms_b200 instead of compiling the
// runFileCopy of src/vFlow.cpp -- or, without touching the reference's build, as a shared library placed in
// front of it: `make -C oracle dropin` in this repository does exactly that with the unmodified sources)
#include "../include/vFlow.h"
#include "farms_b200.h"
#include <chrono>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>

long vFlowManager::runFileCopy(unsigned long int NUMEVENTS)   // same signature as src/vFlow.cpp:111
{
    const std::string outFileName = fileNameInput + "_FARMSOut_batch.txt";     // src/vFlow.cpp:131
    std::ifstream eventsFile((fileNameInput + ".txt").c_str());                // :150-156
    std::string line;
    int x_ = 0, y_ = 0, pol_ = 0;
    unsigned int time_ = 0;
    while (std::getline(eventsFile, line) && X.size() < NUMEVENTS) {           // :173-188
        std::stringstream stream(line);
        stream >> x_ >> y_ >> time_ >> pol_;
        X.push_back(x_); Y.push_back(y_); T.push_back(time_); POL.push_back(pol_);
    }
    const size_t n = X.size();
    std::cout << "Done reading " << n << " Events." << std::endl;
    if (n == 0) throw std::out_of_range("no events");                          // T.at(0), :194
    std::vector<uint16_t> x(n), y(n);
    std::vector<uint64_t> t(n);
    for (size_t i = 0; i < n; i++) { x[i] = X[i]; y[i] = Y[i]; t[i] = T[i]; }

    farms_config cfg = {};
    cfg.width = width; cfg.height = height;          // note: the ctor takes height first (vFlow.h:100)
    cfg.filtersize = 2 * fRad + 1;                   // already normalised; normalising twice is idempotent
    cfg.inlier_check = minEvtsOnPlane;
    cfg.flags = FARMS_FLAG_EXACT_POOLING;            // text output: FP64 sums throughout (drop it for throughput)
    farms_ctx *ctx = nullptr;
    if (farms_create(&ctx, &cfg) != FARMS_OK) throw std::runtime_error("no B200 device");

    std::vector<uint32_t> trel(n);
    std::vector<double> gr(n), gth(n), vx(n), vy(n), lr(n), lth(n);
    std::vector<uint8_t> scale(n);
    farms_out out = {};
    out.t_rel = trel.data(); out.global_r = gr.data(); out.global_theta = gth.data();
    out.vx = vx.data(); out.vy = vy.data(); out.local_r = lr.data(); out.local_theta = lth.data();
    out.scale = scale.data();

    auto a = std::chrono::system_clock::now();
    int rc = farms_process_host(ctx, x.data(), y.data(), t.data(), nullptr, n, &out);
    auto b = std::chrono::system_clock::now();
    if (rc != FARMS_OK) { std::string e = farms_last_error(ctx); farms_destroy(ctx); throw std::runtime_error(e); }
    this->numEvents = (double)farms_num_events(ctx);
    farms_destroy(ctx);

    std::ofstream f(outFileName.c_str());
    for (size_t i = 0; i < n; i++)                    // the reference's row, src/vFlow.cpp:438
        f << X[i] << " " << Y[i] << " " << (int)trel[i] << " " << (POL[i] < 0 ? 0 : POL[i]) << " " << gr[i] << " "
          << gth[i] << " " << vx[i] << " " << vy[i] << " " << lr[i] << " " << lth[i] << " " << (int)scale[i] << std::endl;
    return std::chrono::duration_cast<std::chrono::microseconds>(b - a).count();
}
