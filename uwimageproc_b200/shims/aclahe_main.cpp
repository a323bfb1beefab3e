// aclahe_main.cpp - the aclahe command line of the reference on top of libuwip.so.
//   replaces: modules/aclahe/src/aclahe.cpp:64-226.  The reference converts to HSV, sweeps cv::CLAHE over 5 block sizes x
//   51 clip limits on the V channel, prints the 5 x 51 entropy table (:199-206) and ends at its TODO list (:209-218).
//   Here the table comes from ONE call (uwip_clahe_entropy_sweep_u8_dev: one pixel pass per grid, no images written);
//   `-bs= -cl=` then apply CLAHE with the chosen pair through the frame wrapper (BGR -> HSV, CLAHE on V, HSV -> BGR).
//   The knee search of the Python prototype (scipy curve_fit / splines, ACLAHE.py:69-96) lives in
//   uwimageproc_b200/cli/aclahe.py; this binary takes the pair as arguments.
#include <iostream>
#include <string>
#include <vector>

#include <opencv2/core.hpp>
#ifndef UWIP_OPENCV_STANDIN
#include <opencv2/highgui.hpp>
#include <opencv2/imgcodecs.hpp>
#endif

#include "../../include/uwip.h"

int main(int argc, char* argv[]) {
  cv::String keys =
      "{@input |<none>  | Input video path}"
      "{@output |<none> | Prefix for output file}"
      "{bs      |       | block size to apply (with cl)}"
      "{cl      |       | clip limit to apply (with bs)}"
      "{help h usage ?  |       | show this help message}";
  cv::CommandLineParser cvParser(argc, argv, keys);
  std::cout << "ACLAHE: a C++ implementation Automatic Contrast Limited Adaptive Histogram Equalization" << std::endl;
  std::cout << "Built with OpenCV " << CV_VERSION << std::endl;
  if (argc < 3 || cvParser.has("help")) {
    std::cout << std::endl << "\tExample:" << std::endl;
    std::cout << "\t$ aclahe input.jpg output.jpg" << std::endl;
    std::cout << "\tThis will apply ACLAHE to gray levels of 'input.jpg' image file, and save it into 'output.jg'" << std::endl << std::endl;
    return 0;
  }
  cv::String InputFile = cvParser.get<cv::String>(0), OutputFile = cvParser.get<cv::String>(1);
  std::cout << "***************************************" << std::endl;
  std::cout << "Input: " << InputFile << std::endl;
  std::cout << "Output: " << OutputFile << std::endl;
  cv::Mat src = cv::imread(InputFile, cv::IMREAD_COLOR);
  if (src.empty()) { std::cout << "Failed to read input image, exiting..." << std::endl; return -1; }
  std::cout << "Input image loaded..." << std::endl;
  uwip_ctx* ctx = nullptr;
  if (uwip_create(0, &ctx) != UWIP_OK) { std::cout << "No CUDA device detected" << std::endl; return -1; }
  // V = max(B, G, R): plane 2 of cvtColor(BGR2HSV) (aclahe.cpp:152-154)
  const int w = src.cols, h = src.rows;
  std::vector<unsigned char> v((size_t)w * h);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const unsigned char* p = src.data + (size_t)y * src.step + 3 * x;
      unsigned char m = p[0] > p[1] ? p[0] : p[1];
      v[(size_t)y * w + x] = m > p[2] ? m : p[2];
    }
  const int BlockSize[5] = {2, 4, 8, 16, 32};
  std::vector<double> clips;
  for (float cl = 0.0f; cl <= 25.0f; cl += 0.5f) clips.push_back(cl);   // aclahe.cpp:160-163,180
  std::vector<float> ent(5 * clips.size());
  void* d_v = nullptr;
  int rc = uwip_device_alloc(ctx, v.size(), &d_v);
  if (rc == UWIP_OK) rc = uwip_copy_h2d(ctx, d_v, v.data(), v.size());
  if (rc == UWIP_OK)
    rc = uwip_clahe_entropy_sweep_u8_dev(ctx, (const uint8_t*)d_v, 1, w, h, BlockSize, 5, clips.data(), (int)clips.size(), 0 /* entropy flavour of aclaheEntropy, aclahe.cpp:228-248 */, ent.data());
  if (rc != UWIP_OK) { std::cout << "sweep failed: " << uwip_last_error(ctx) << std::endl; uwip_destroy(ctx); return -1; }
  for (int i = 0; i < 5; i++) {   // the table of aclahe.cpp:199-206
    for (size_t j = 0; j < clips.size(); j++) std::cout << ent[i * clips.size() + j] << " ";
    std::cout << std::endl;
  }
  if (cvParser.has("bs") && cvParser.has("cl")) {
    const int bs = cvParser.get<int>("bs");
    const double cl = cvParser.get<double>("cl");
    cv::Mat dst(h, w, CV_8UC3);
    rc = uwip_aclahe_bgr8(ctx, src.data, (size_t)src.step, dst.data, (size_t)dst.step, w, h, cl, bs, bs, UWIP_HSV_ROUND_CV2_4_13);
    if (rc != UWIP_OK) { std::cout << "aclahe failed: " << uwip_last_error(ctx) << std::endl; uwip_destroy(ctx); return -1; }
    cv::imwrite(OutputFile, dst);
    std::cout << "saved " << OutputFile << " (BS = " << bs << ", CL = " << cl << ")" << std::endl;
  }
  uwip_device_free(ctx, d_v);
  uwip_destroy(ctx);
  return 0;
}
