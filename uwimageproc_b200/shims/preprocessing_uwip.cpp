// preprocessing_uwip.cpp - modules/common/preprocessing.cpp re-implemented on the C ABI of libuwip.so.
// Same names, argument meaning and in-place behaviour as the reference (file:line cited per function);
// no pixel arithmetic happens on the host and there is no CPU fallback: when libuwip cannot create a
// context (no sm_100 device) or a call fails, the error is printed and the process aborts; with
// UWIP_SHIM_SOFT_ERRORS=1 the call records the error and leaves its outputs untouched instead.
#include "preprocessing.h"

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/uwip.h"

namespace {
struct ShimState {
  uwip_ctx* ctx = nullptr;
  int status = 0;
  std::string err;
  ~ShimState() { if (ctx) uwip_destroy(ctx); }
};
thread_local ShimState g_shim;  // one context per host thread (the C ABI's threading contract)

uwip_ctx* shim_ctx() {
  if (!g_shim.ctx) {
    int rc = uwip_create(0, &g_shim.ctx);  // the reference always uses device 0 (histretch.cpp:134)
    if (rc != UWIP_OK) {
      g_shim.status = rc;
      g_shim.err = uwip_last_error(nullptr);
      std::fprintf(stderr, "uwip shim: %s\n", g_shim.err.c_str());
      g_shim.ctx = nullptr;
      const char* soft = std::getenv("UWIP_SHIM_SOFT_ERRORS");
      if (!(soft && soft[0] == '1')) std::abort();
    }
  }
  return g_shim.ctx;
}
// The reference functions are void: a failure cannot be returned.  It is LOUD by default - the message goes to stderr and the
// process aborts (a frame silently left unstretched is worse than a crash) - unless UWIP_SHIM_SOFT_ERRORS=1 asks for the
// "print and continue" behaviour of the reference (histretch.cpp:252); the status stays readable through uwipShimLastStatus().
void shim_done(int rc) {
  g_shim.status = rc;
  if (rc != UWIP_OK) {
    g_shim.err = g_shim.ctx ? uwip_last_error(g_shim.ctx) : "no usable CUDA device (libuwip has no CPU fallback)";
    std::fprintf(stderr, "uwip shim: %s\n", g_shim.err.c_str());
    const char* soft = std::getenv("UWIP_SHIM_SOFT_ERRORS");
    if (!(soft && soft[0] == '1')) std::abort();
  }
}
bool is_u8_plane(const cv::Mat& m) { return !m.empty() && m.type() == CV_8UC1; }
}  // namespace

int uwipShimLastStatus() { return g_shim.status; }
const char* uwipShimLastError() { return g_shim.err.c_str(); }

void getHistogram(cv::Mat* img, cv::Mat* dstHist) {
  uwip_ctx* ctx = shim_ctx();
  if (!ctx || !img || !dstHist || !is_u8_plane(*img)) { if (ctx) shim_done(UWIP_ERR_INVALID); return; }
  dstHist->create(256, 1, CV_32FC1);
  float hist[256];
  int rc = uwip_histogram_u8(ctx, img->data, img->cols, img->rows, (size_t)img->step, hist);
  if (rc == UWIP_OK)
    for (int i = 0; i < 256; i++) dstHist->at<float>(i, 0) = hist[i];
  shim_done(rc);
}

void imgChannelStretch(cv::Mat imgOriginal, cv::Mat imgStretched, int lowerPercentile, int higherPercentile) {
  uwip_ctx* ctx = shim_ctx();
  if (!ctx) return;
  if (!is_u8_plane(imgOriginal) || !is_u8_plane(imgStretched) || imgOriginal.rows != imgStretched.rows || imgOriginal.cols != imgStretched.cols) {
    shim_done(UWIP_ERR_INVALID);  // preprocessing.h:60-64: same dimensions required
    return;
  }
  // histogram of imgOriginal, arithmetic written to imgStretched (in place when they share pixels,
  // which is what every reference caller does: histretch.cpp:236,247)
  shim_done(uwip_channel_stretch_u8(ctx, imgOriginal.data, (size_t)imgOriginal.step, imgStretched.data, (size_t)imgStretched.step, imgOriginal.cols,
                                    imgOriginal.rows, lowerPercentile, higherPercentile, nullptr, nullptr));
}

#if USE_GPU
void imgChannelStretchGPU(cv::cuda::GpuMat imgOriginal, cv::cuda::GpuMat imgStretched, int lowerPercentile, int higherPercentile) {
  uwip_ctx* ctx = shim_ctx();
  if (!ctx) return;
  if (imgOriginal.rows != imgStretched.rows || imgOriginal.cols != imgStretched.cols) {
    shim_done(UWIP_ERR_INVALID);  // preprocessing.h:60-64: same dimensions required
    return;
  }
  // GpuMat planes are pitched (cudaMallocPitch): the pitched entry point takes them as they are
  int rc = uwip_channel_stretch_u8_dev_pitched(ctx, imgOriginal.data, (size_t)imgOriginal.step, imgStretched.data, (size_t)imgStretched.step,
                                               imgOriginal.cols, imgOriginal.rows, lowerPercentile, higherPercentile);
  if (rc == UWIP_OK) rc = uwip_synchronize(ctx);  // cv::cuda::add / multiply of the reference block too
  shim_done(rc);
}
#endif

int numChannel(char c) { return uwip_num_channel(c); }
int numSpace(char c) { return uwip_num_space(c); }

float aclaheEntropy(cv::Mat img) {
  uwip_ctx* ctx = shim_ctx();
  float e = 0.f;
  if (!ctx) return e;
  if (!is_u8_plane(img)) { shim_done(UWIP_ERR_INVALID); return e; }
  shim_done(uwip_entropy_u8(ctx, img.data, img.cols, img.rows, (size_t)img.step, /*flavour: aclahe.cpp*/ 0, &e));
  return e;
}

// videostrip.cpp:170-184 (calcBlur) and :39-60 (calcBlurGPU): BGR frame in, float stdev out
static float calc_blur_shim(const cv::Mat& frame, int aperture) {
  uwip_ctx* ctx = shim_ctx();
  float sd = 0.f;
  if (!ctx) return sd;
  if (frame.empty() || frame.type() != CV_8UC3) { shim_done(UWIP_ERR_INVALID); return sd; }
  shim_done(uwip_calc_blur_bgr8(ctx, frame.data, (size_t)frame.step, frame.cols, frame.rows, aperture, &sd, nullptr, nullptr, 0));
  return sd;
}
float calcBlur(cv::Mat frame) { return calc_blur_shim(frame, 3); }
float calcBlurGPU(cv::Mat frame) { return calc_blur_shim(frame, 1); }
