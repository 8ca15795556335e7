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

// Stages of the coder on their own, for direct comparison with FSE_normalizeCount / FSE_writeNCount / FSE_readNCount of the reference library
// on inputs that FSE_compress itself rarely produces (fallback-normalisation branches, long zero runs in the header).
extern "C" int fse_host_normalize(int16_t* norm, uint32_t table_log, const uint32_t* count, uint32_t total, uint32_t max_symbol) {
    return fse::normalize(norm, table_log, count, total, max_symbol);
}

extern "C" int fse_host_write_ncount(uint8_t* out, uint32_t cap, const int16_t* norm, uint32_t max_symbol, uint32_t table_log) {
    return fse::write_ncount(out, cap, norm, max_symbol, table_log);
}

extern "C" int fse_host_read_ncount(int16_t* norm, uint32_t* max_symbol, uint32_t* table_log, const uint8_t* hdr, uint32_t hdr_size) {
    return fse::read_ncount(norm, max_symbol, table_log, hdr, hdr_size);
}
