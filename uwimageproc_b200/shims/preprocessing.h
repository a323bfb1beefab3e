// Drop-in replacement for modules/common/preprocessing.h of MecatronicaUSB/uwimageproc: the same
// global-namespace signatures (preprocessing.h:38,66,109,112,115, aclahe.cpp:58 and videostrip.hpp:98,106), implemented on
// libuwip.so (include/uwip.h) instead of OpenCV's CPU kernels.  A module that includes this header and
// links preprocessing_uwip.cpp + libuwip.so in place of ../common/preprocessing.cpp needs no other change.
#ifndef UWIP_SHIM_PREPROCESSING_H
#define UWIP_SHIM_PREPROCESSING_H

#include <opencv2/core.hpp>

// void getHistogram(cv::Mat *img, cv::Mat *dstHist)   (preprocessing.h:38, preprocessing.cpp:25-34)
// allocates *dstHist as 256x1 CV_32F like calcHist does.
void getHistogram(cv::Mat* img, cv::Mat* dstHist);

// void imgChannelStretch(cv::Mat, cv::Mat, int=0, int=100)   (preprocessing.h:66, preprocessing.cpp:74-105)
// Both Mats are header copies sharing the caller's pixels; imgStretched is modified in place.
void imgChannelStretch(cv::Mat imgOriginal, cv::Mat imgStretched, int lowerPercentile = 0, int higherPercentile = 100);

#if USE_GPU
// void imgChannelStretchGPU(cv::cuda::GpuMat, cv::cuda::GpuMat, int, int)   (preprocessing.h:109,
// preprocessing.cpp:109-144): the planes already live on the device; no download of the plane.
void imgChannelStretchGPU(cv::cuda::GpuMat imgOriginal, cv::cuda::GpuMat imgStretched, int lowerPercentile, int higherPercentile);
#endif

int numChannel(char c);  // preprocessing.h:112
int numSpace(char c);    // preprocessing.h:115

// float aclaheEntropy(cv::Mat img)   (aclahe.cpp:58, 228-248)
float aclaheEntropy(cv::Mat img);

// float calcBlur(Mat frame) / float calcBlurGPU(Mat frame)   (modules/videostrip/include/videostrip.hpp:98,106;
// videostrip.cpp:170-184, 39-60): standard deviation of the 8-bit Laplacian of the grey frame.  calcBlur's call
// `Laplacian(grey, laplacian, grey.type(), CV_16S)` means aperture 3; calcBlurGPU asks cv::cuda for aperture 1.
float calcBlur(cv::Mat frame);
float calcBlurGPU(cv::Mat frame);

// What the reference's "print and continue" becomes across an ABI: the status of the last shim call on
// this thread (0 = UWIP_OK) and its text.  The reference signatures themselves stay void.
int uwipShimLastStatus();
const char* uwipShimLastError();

#endif
