// JPEG decode in front of the clip loader ("nvJPEG decode is the next wall"): the reference decodes every
// frame on the host with PIL (mega_core/data/datasets/vid.py `Image.open(...).convert("RGB")`) and ships fp32 pixels; here
// the compressed bytes cross PCIe and the frame is decoded on the GPU straight into the uint8 HWC layout that
// dvid_resize_bilinear_u8 (the byte-exact Pillow resize) consumes.
//
// This is a LIBRARY stage, not a hand-written kernel: nvJPEG (CUDA toolkit) does Huffman decoding + IDCT + colour
// conversion.  The library is opened lazily with dlopen, so libdvid_b200.so has no link-time dependency on it and the
// detector works on machines without it (the two entry points then return DVID_ERR_DRIVER).
#include <dlfcn.h>
#include <nvjpeg.h>
#include "dvid_internal.h"

namespace dvid {

namespace {

struct NvJpeg {
  void* lib = nullptr;
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t state = nullptr;
  decltype(&nvjpegCreateSimple) create = nullptr;
  decltype(&nvjpegJpegStateCreate) state_create = nullptr;
  decltype(&nvjpegGetImageInfo) info = nullptr;
  decltype(&nvjpegDecode) decode = nullptr;
  bool tried = false, ok = false;
};

NvJpeg& nvj() {
  static NvJpeg j;
  if (j.tried) return j;
  j.tried = true;
  for (const char* name : {"libnvjpeg.so.12", "libnvjpeg.so"}) {
    j.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
    if (j.lib) break;
  }
  if (!j.lib) return j;
  j.create = reinterpret_cast<decltype(j.create)>(dlsym(j.lib, "nvjpegCreateSimple"));
  j.state_create = reinterpret_cast<decltype(j.state_create)>(dlsym(j.lib, "nvjpegJpegStateCreate"));
  j.info = reinterpret_cast<decltype(j.info)>(dlsym(j.lib, "nvjpegGetImageInfo"));
  j.decode = reinterpret_cast<decltype(j.decode)>(dlsym(j.lib, "nvjpegDecode"));
  if (!j.create || !j.state_create || !j.info || !j.decode) return j;
  if (j.create(&j.handle) != NVJPEG_STATUS_SUCCESS) return j;
  if (j.state_create(j.handle, &j.state) != NVJPEG_STATUS_SUCCESS) return j;
  j.ok = true;
  return j;
}

}  // namespace

int jpeg_info(const unsigned char* data, long nbytes, int* width, int* height) {
  NvJpeg& j = nvj();
  if (!j.ok) return DVID_ERR_DRIVER;
  int comps = 0;
  nvjpegChromaSubsampling_t ss;
  int ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
  if (j.info(j.handle, data, static_cast<size_t>(nbytes), &comps, &ss, ws, hs) != NVJPEG_STATUS_SUCCESS)
    return DVID_ERR_ARG;
  *width = ws[0];
  *height = hs[0];
  return DVID_OK;
}

int jpeg_decode_rgb(const unsigned char* data, long nbytes, unsigned char* dst_hwc, int width, int height,
                    cudaStream_t stream) {
  NvJpeg& j = nvj();
  if (!j.ok) return DVID_ERR_DRIVER;
  nvjpegImage_t img = {};
  img.channel[0] = dst_hwc;
  img.pitch[0] = static_cast<size_t>(width) * 3;
  const nvjpegStatus_t st = j.decode(j.handle, j.state, data, static_cast<size_t>(nbytes), NVJPEG_OUTPUT_RGBI, &img, stream);
  return st == NVJPEG_STATUS_SUCCESS ? DVID_OK : DVID_ERR_CUDA;
}

}  // namespace dvid
