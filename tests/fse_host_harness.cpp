// TEST INFRASTRUCTURE: compiles the product's host/device FSE core (neural_imaging_b200/csrc/fse_core.cuh) with g++ so that the exact
// source the CUDA kernels instantiate can be compared byte for byte with the reference library (oracle/_ref/libfse_ref.so) and with the
// committed golden vectors on a machine without a GPU. Never loaded by the product (there is no CPU fallback).
#include "../neural_imaging_b200/csrc/fse_core.cuh"

extern "C" int fse_host_compress(uint8_t* dst, uint32_t cap, const uint8_t* src, uint32_t n) {
    static fse::EncScratch S;
    fse::histogram(src, n, S.count);
    return fse::compress_counted(dst, cap, src, n, S);
}

extern "C" int fse_host_decompress(uint8_t* dst, uint32_t cap, const uint8_t* src, uint32_t n) {
    static fse::DecScratch S;
    return fse::decompress(dst, cap, src, n, S);
}
