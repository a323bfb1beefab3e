// jpegio.cu - JPEG files straight to and from the device: nvJPEG decodes into the bgr8 layout the chain
// reads and encodes from the one it writes, so a real (non-synthetic) frame crosses PCIe as its compressed bytes only.
//   replaces: imread / imwrite around the chain - modules/histretch/src/histretch.cpp:158,268, modules/aclahe/src/aclahe.cpp:135,
//             modules/bgdehaze/main.py:16,19 - for baseline JPEG files.
// nvJPEG is library code (like cuBLAS for a GEMM): file decoding is not on the hot path of BASELINE.json; the pixels that come
// out of its IDCT are NOT bit-identical to libjpeg-turbo's (cv2.imread), so parity is stated on the decoded pixels
// (tests/test_gpu_modules.py::test_jpeg_device_io: decode within a few levels of cv2.imdecode - the 4:2:0 chroma upsampling differs -,
// chain output against the oracle on the pixels nvJPEG produced).
#include <nvjpeg.h>

#include "common.cuh"

struct JpegIo {
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t dec = nullptr;
  nvjpegEncoderState_t enc = nullptr;
  nvjpegEncoderParams_t params = nullptr;
};

static int jpeg_get(uwip_ctx* ctx, JpegIo** out) {
  if (!ctx->jpeg) {
    JpegIo* j = new JpegIo();
    if (nvjpegCreateSimple(&j->handle) != NVJPEG_STATUS_SUCCESS || nvjpegJpegStateCreate(j->handle, &j->dec) != NVJPEG_STATUS_SUCCESS ||
        nvjpegEncoderStateCreate(j->handle, &j->enc, ctx->stream) != NVJPEG_STATUS_SUCCESS ||
        nvjpegEncoderParamsCreate(j->handle, &j->params, ctx->stream) != NVJPEG_STATUS_SUCCESS) {
      delete j;
      uwip_set_err(ctx, "nvJPEG initialisation failed");
      return UWIP_ERR_CUDA;
    }
    ctx->jpeg = j;
  }
  *out = (JpegIo*)ctx->jpeg;
  return UWIP_OK;
}

void jpeg_io_destroy(uwip_ctx* ctx) {
  JpegIo* j = (JpegIo*)ctx->jpeg;
  if (!j) return;
  if (j->params) nvjpegEncoderParamsDestroy(j->params);
  if (j->enc) nvjpegEncoderStateDestroy(j->enc);
  if (j->dec) nvjpegJpegStateDestroy(j->dec);
  if (j->handle) nvjpegDestroy(j->handle);
  delete j;
  ctx->jpeg = nullptr;
}

int jpeg_info(uwip_ctx* ctx, const uint8_t* data, size_t len, int* w, int* h) {
  JpegIo* j;
  UWIP_CHECK(jpeg_get(ctx, &j));
  int nc = 0, ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
  nvjpegChromaSubsampling_t ss;
  if (nvjpegGetImageInfo(j->handle, data, len, &nc, &ss, ws, hs) != NVJPEG_STATUS_SUCCESS) {
    uwip_set_err(ctx, "not a JPEG stream nvJPEG can parse");
    return UWIP_ERR_INVALID;
  }
  *w = ws[0]; *h = hs[0];
  return UWIP_OK;
}

// decode into interleaved B,G,R bytes at d_bgr (w*h*3, contiguous): what imread(..., IMREAD_COLOR) hands the reference
int jpeg_decode_dev(uwip_ctx* ctx, const uint8_t* data, size_t len, uint8_t* d_bgr, int w, int h) {
  JpegIo* j;
  UWIP_CHECK(jpeg_get(ctx, &j));
  nvjpegImage_t img;
  memset(&img, 0, sizeof(img));
  img.channel[0] = d_bgr;
  img.pitch[0] = (size_t)w * 3;
  nvjpegStatus_t st = nvjpegDecode(j->handle, j->dec, data, len, NVJPEG_OUTPUT_BGRI, &img, ctx->stream);
  if (st != NVJPEG_STATUS_SUCCESS) {
    uwip_set_err(ctx, "nvjpegDecode failed (status %d)", (int)st);
    return UWIP_ERR_CUDA;
  }
  (void)h;
  return UWIP_OK;
}

// encode interleaved B,G,R device bytes; 4:2:0 and quality 95 are what cv2.imwrite does by default
int jpeg_encode_dev(uwip_ctx* ctx, const uint8_t* d_bgr, int w, int h, int quality, uint8_t* out, size_t cap, size_t* out_len) {
  JpegIo* j;
  UWIP_CHECK(jpeg_get(ctx, &j));
  nvjpegEncoderParamsSetQuality(j->params, quality, ctx->stream);
  nvjpegEncoderParamsSetSamplingFactors(j->params, NVJPEG_CSS_420, ctx->stream);
  nvjpegEncoderParamsSetOptimizedHuffman(j->params, 0, ctx->stream);
  nvjpegImage_t img;
  memset(&img, 0, sizeof(img));
  img.channel[0] = const_cast<uint8_t*>(d_bgr);
  img.pitch[0] = (size_t)w * 3;
  nvjpegStatus_t st = nvjpegEncodeImage(j->handle, j->enc, j->params, &img, NVJPEG_INPUT_BGRI, w, h, ctx->stream);
  if (st != NVJPEG_STATUS_SUCCESS) {
    uwip_set_err(ctx, "nvjpegEncodeImage failed (status %d)", (int)st);
    return UWIP_ERR_CUDA;
  }
  size_t need = 0;
  if (nvjpegEncodeRetrieveBitstream(j->handle, j->enc, nullptr, &need, ctx->stream) != NVJPEG_STATUS_SUCCESS) {
    uwip_set_err(ctx, "nvjpegEncodeRetrieveBitstream (size) failed");
    return UWIP_ERR_CUDA;
  }
  *out_len = need;
  if (!out || cap < need) {
    if (out) uwip_set_err(ctx, "output buffer too small for the JPEG stream (%zu bytes needed)", need);
    return out ? UWIP_ERR_INVALID : UWIP_OK;   // out == NULL: size query
  }
  UWIP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (nvjpegEncodeRetrieveBitstream(j->handle, j->enc, out, &need, ctx->stream) != NVJPEG_STATUS_SUCCESS) {
    uwip_set_err(ctx, "nvjpegEncodeRetrieveBitstream failed");
    return UWIP_ERR_CUDA;
  }
  UWIP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return UWIP_OK;
}
