// histretch_main.cpp - the histretch command line of the reference on top of libuwip.so.
//   replaces: modules/histretch/src/histretch.cpp:61-271 (same positional arguments, the same -c / -cuda / -time keys of
//   the cv::CommandLineParser string at :68-74, the same console messages).  The colour conversions, split / merge and
//   imgChannelStretch calls of the channel loop (:219-254) become ONE call, uwip_histretch_bgr8, which takes the letter
//   string.  HighGUI windows (imshow / waitKey, :159-160,263-270) are not opened.
// Build: g++ -std=c++11 histretch_main.cpp preprocessing_uwip.cpp $(pkg-config --cflags --libs opencv4) -L.. -luwip
//        (check.sh builds it against the stand-in header, whose imread / imwrite speak PPM).
#include <iostream>
#include <string>

#include <opencv2/core.hpp>
#ifndef UWIP_OPENCV_STANDIN
#include <opencv2/highgui.hpp>
#include <opencv2/imgcodecs.hpp>
#endif

#include "../../include/uwip.h"
#include "preprocessing.h"

#define ABOUT_STRING "Histogram Stretching tool (B200 build)"

int main(int argc, char* argv[]) {
  cv::String keys =
      "{@input |<none>  | Input image file}"
      "{@output |<none> | Output image file}"
      "{c       |r      | Channel to apply histogram equalization}"
      "{cuda    |       | Use CUDA or not (CUDA ON: 1, CUDA OFF: 0)}"
      "{time    |       | Show time measurements or not (ON: 1, OFF: 0)}"
      "{literal |       | 1: reproduce the channel loop as written (histretch.cpp:238-240), 0: the intended order}"
      "{help h usage ?  |       | show this help message}";
  cv::CommandLineParser cvParser(argc, argv, keys);
  cvParser.about(ABOUT_STRING);
  std::cout << ABOUT_STRING << std::endl;
  std::cout << "Built with OpenCV " << CV_VERSION << std::endl;
  if (argc < 3 || cvParser.has("help")) {
    std::cout << "C++ implementation of Histogram Stretching for specific channels of input image" << std::endl;
    cvParser.printMessage();
    std::cout << "Argument 'c=<channels>' is a string containing an ordered list of desired channels to be stretched" << std::endl;
    std::cout << "\t-c=R|G|B\tfor RGB space\n\t-c=H|S|V\tfor HSV space\n\t-c=h|s|l\tfor HSL space\n\t-c=L|a|b\tfor Lab space\n"
                 "\t-c=Y|C|X\tfor YCrCb space\n\t-cuda=0 or -cuda=1 (CUDA ON: 1, CUDA OFF: 0, if available)" << std::endl;
    std::cout << "\n\tExample:\n\t$ histretch -c=HV input.jpg output.jpg -cuda=0 -time=1" << std::endl;
    return 0;
  }
  cv::String InputFile = cvParser.get<cv::String>(0), OutputFile = cvParser.get<cv::String>(1);
  cv::String cChannel = cvParser.has("c") ? cvParser.get<cv::String>("c") : cv::String("r");   // default of the key string
  int Time = cvParser.get<int>("time");
  int literal = cvParser.get<int>("literal");
  if (!cvParser.check()) { cvParser.printErrors(); return -1; }
  if (cvParser.has("cuda") && cvParser.get<int>("cuda") == 0)
    std::cout << "CUDA deactivated: this build has no CPU path, running on the GPU" << std::endl;
  std::cout << "***************************************" << std::endl;
  std::cout << "Input: " << InputFile << std::endl;
  std::cout << "Output: " << OutputFile << std::endl;
  std::cout << "Channel: " << cChannel << std::endl;
  cv::Mat src = cv::imread(InputFile, cv::IMREAD_COLOR);
  if (src.empty()) { std::cout << "Failed to read input image, exiting..." << std::endl; return -1; }
  const int num_convert = (int)cChannel.length();
  const int min_percent = 2, max_percent = 98;   // histretch.cpp:154
  std::cout << "Applying " << num_convert << " histretch" << std::endl;
  uwip_ctx* ctx = nullptr;
  if (uwip_create(0, &ctx) != UWIP_OK) { std::cout << "No CUDA device detected" << std::endl; return -1; }
  double t = (double)cv::getTickCount();
  for (int nc = 0; nc < num_convert; nc++) {
    char c = cChannel[nc];
    std::cout << "\tChannel[" << nc << "]: " << c << std::endl;
    if (numSpace(c) == -1) std::cout << "Option " << c << " not recognized, skipping..." << std::endl;
  }
  cv::Mat dst(src.rows, src.cols, CV_8UC3);
  int rc = uwip_histretch_bgr8(ctx, src.data, (size_t)src.step, dst.data, (size_t)dst.step, src.cols, src.rows, cChannel.c_str(), min_percent,
                               max_percent, literal ? UWIP_ORDER_LITERAL : UWIP_ORDER_INTENDED, UWIP_HSV_ROUND_CV2_4_13);
  if (rc != UWIP_OK) { std::cout << "histretch failed: " << uwip_last_error(ctx) << std::endl; uwip_destroy(ctx); return -1; }
  if (Time == 1) {
    t = 1000 * ((double)cv::getTickCount() - t) / cv::getTickFrequency();
    std::cout << std::endl << "Execution Time GPU :" << t << " ms " << std::endl;
  }
  std::cout << "hS: saving to disk" << std::endl;
  cv::imwrite(OutputFile, dst);
  uwip_destroy(ctx);
  return 0;
}
