// JPEG decode stage in front of the resize kernel: the reference decodes on the host
// (torchvision.io.read_image -> PIL, DataRetrieval.py:70-96, 175-180); here the entropy-coded stream is handed to nvJPEG
// (library call: Huffman decode + IDCT; default hybrid backend) and the 8-bit luma plane lands in device memory, where
// bv_resize_center_crop_u8 and the stem kernel take over - decoded pixels never exist in host memory.
//
// nvJPEG is resolved with dlopen at first use, so libbiovil_b200.so loads (and every other entry point works) on a
// machine without libnvjpeg.  One nvjpeg handle + decoder state per device and host thread.
// Parity: JPEG decoders differ in their IDCT rounding; against libjpeg-turbo (Pillow, the reference's decoder) the luma
// plane is within +-2 grey levels, tested in tests/test_jpeg_gpu.py - this stage is NOT bit-exact, unlike the resize.
#pragma once
#include <dlfcn.h>
#include <nvjpeg.h>

namespace jpeg_stage {

struct Api {
    void* lib = nullptr;
    nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t*) = nullptr;
    nvjpegStatus_t (*Destroy)(nvjpegHandle_t) = nullptr;
    nvjpegStatus_t (*JpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
    nvjpegStatus_t (*JpegStateDestroy)(nvjpegJpegState_t) = nullptr;
    nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*,
                                   int*) = nullptr;
    nvjpegStatus_t (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t,
                             nvjpegImage_t*, cudaStream_t) = nullptr;
    // batched decode (GPU-assisted Huffman / hardware JPEG engines); optional: older libraries may lack them
    nvjpegStatus_t (*CreateEx)(nvjpegBackend_t, nvjpegDevAllocator_t*, nvjpegPinnedAllocator_t*, unsigned int, nvjpegHandle_t*) = nullptr;
    nvjpegStatus_t (*BatchedInitialize)(nvjpegHandle_t, nvjpegJpegState_t, int, int, nvjpegOutputFormat_t) = nullptr;
    nvjpegStatus_t (*Batched)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char* const*, const size_t*, nvjpegImage_t*,
                              cudaStream_t) = nullptr;
};

inline const Api* api(const char** why) {
    static Api a;
    static bool tried = false;
    static const char* err = nullptr;
    if (!tried) {
        tried = true;
        for (const char* name : {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"}) {
            a.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (a.lib) break;
        }
        if (!a.lib) {
            err = "libnvjpeg.so.12 not found (dlopen)";
        } else {
#define JS_SYM(field, sym)                                                  \
    a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.lib, sym));       \
    if (!a.field) err = "nvJPEG symbol missing: " sym;
            JS_SYM(CreateSimple, "nvjpegCreateSimple")
            JS_SYM(Destroy, "nvjpegDestroy")
            JS_SYM(JpegStateCreate, "nvjpegJpegStateCreate")
            JS_SYM(JpegStateDestroy, "nvjpegJpegStateDestroy")
            JS_SYM(GetImageInfo, "nvjpegGetImageInfo")
            JS_SYM(Decode, "nvjpegDecode")
#undef JS_SYM
            a.CreateEx = reinterpret_cast<decltype(a.CreateEx)>(dlsym(a.lib, "nvjpegCreateEx"));
            a.BatchedInitialize = reinterpret_cast<decltype(a.BatchedInitialize)>(dlsym(a.lib, "nvjpegDecodeBatchedInitialize"));
            a.Batched = reinterpret_cast<decltype(a.Batched)>(dlsym(a.lib, "nvjpegDecodeBatched"));
        }
    }
    if (why) *why = err;
    return err ? nullptr : &a;
}

struct Decoder {
    nvjpegHandle_t handle = nullptr;
    nvjpegJpegState_t state = nullptr;
};

// decoder of the calling thread for the current device (created on first use; lives for the process)
inline Decoder* decoder(int device, const Api* a, const char** why) {
    thread_local Decoder dec[64];
    if (device < 0 || device >= 64) {
        *why = "device index out of range";
        return nullptr;
    }
    Decoder* d = &dec[device];
    if (!d->handle) {
        if (a->CreateSimple(&d->handle) != NVJPEG_STATUS_SUCCESS) {
            d->handle = nullptr;
            *why = "nvjpegCreateSimple failed";
            return nullptr;
        }
        if (a->JpegStateCreate(d->handle, &d->state) != NVJPEG_STATUS_SUCCESS) {
            a->Destroy(d->handle);
            d->handle = nullptr;
            *why = "nvjpegJpegStateCreate failed";
            return nullptr;
        }
    }
    return d;
}

// Batched decoders, one per (host thread, device, backend): several host threads may each push their own sub-batch
// (own nvJPEG handle + state, own stream), which is how the GPU-assisted Huffman backend is kept fed.
struct BatchDecoder {
    nvjpegHandle_t handle = nullptr;
    nvjpegJpegState_t state = nullptr;
    int tried = 0;       // 0 = not yet, 1 = available, -1 = this backend cannot be created on this device / library
    int batch = 0;       // batch size the state was initialised for
};

inline BatchDecoder* batch_decoder(int device, int backend, const Api* a) {
    thread_local BatchDecoder dec[64][8];
    if (device < 0 || device >= 64 || backend < 0 || backend >= 8) return nullptr;
    BatchDecoder* d = &dec[device][backend];
    if (d->tried == 0) {
        d->tried = -1;
        if (a->CreateEx && a->BatchedInitialize && a->Batched &&
            a->CreateEx(static_cast<nvjpegBackend_t>(backend), nullptr, nullptr, 0, &d->handle) == NVJPEG_STATUS_SUCCESS) {
            if (a->JpegStateCreate(d->handle, &d->state) == NVJPEG_STATUS_SUCCESS) d->tried = 1;
            else a->Destroy(d->handle);
        }
    }
    return d->tried == 1 ? d : nullptr;
}

}  // namespace jpeg_stage
