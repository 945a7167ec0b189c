// jpegdecode_files.cpp — batched decode of a directory through the file-ingestion extension.
//
// This repository's own sample (the reference's three samples are built UNMODIFIED from /root/reference/samples by the
// same Makefile). It follows the reference's jpegDecodeBatched sample (samples/jpegDecodeBatched/jpegdecodebatched.cpp:
// list the files, create batch_size stream handles, per batch: read + parse, rocJpegGetImageInfo, size and allocate the
// outputs, rocJpegDecodeBatched, report images/s and MP/s) with one difference: instead of an ifstream read per image into
// a pageable vector on the decode thread (samples/rocjpeg_samples_utils.h:213-234) followed by rocJpegStreamParse, the
// files are read by I/O threads straight into the stream handles' pooled page-locked memory and parsed there
// (rocJpegB200StreamLoadFiles), and the timer covers read + parse + decode.
//
//   jpegdecode_files -i <directory or file> [-fmt native|yuv_planar|y|rgb|rgb_planar] [-b batch] [-t io_threads] [-d device] [-n passes]
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <string>
#include <vector>

#include "rocjpeg.h"
#include "rocjpeg_b200_ext.h"

#define CHECK(call)                                                                                   \
    do {                                                                                              \
        RocJpegStatus st_ = (call);                                                                   \
        if (st_ != ROCJPEG_STATUS_SUCCESS) {                                                          \
            std::fprintf(stderr, "%s returned %s at %s:%d\n", #call, rocJpegGetErrorName(st_), __FILE__, __LINE__); \
            return 1;                                                                                 \
        }                                                                                             \
    } while (0)

// channel sizes as the reference's samples size them (samples/rocjpeg_samples_utils.h:318-399)
static int ChannelSizes(RocJpegOutputFormat fmt, RocJpegChromaSubsampling css, const uint32_t* w, const uint32_t* h, RocJpegImage* img, size_t* bytes) {
    std::memset(img, 0, sizeof(*img));
    bytes[0] = bytes[1] = bytes[2] = 0;
    const uint32_t W = w[0], H = h[0];
    switch (fmt) {
        case ROCJPEG_OUTPUT_RGB: img->pitch[0] = 3 * W; bytes[0] = size_t(3) * W * H; return 1;
        case ROCJPEG_OUTPUT_RGB_PLANAR: for (int c = 0; c < 3; c++) { img->pitch[c] = W; bytes[c] = size_t(W) * H; } return 3;
        case ROCJPEG_OUTPUT_Y: img->pitch[0] = W; bytes[0] = size_t(W) * H; return 1;
        case ROCJPEG_OUTPUT_YUV_PLANAR:
            img->pitch[0] = W; bytes[0] = size_t(W) * H;
            if (css == ROCJPEG_CSS_400) return 1;
            for (int c = 1; c < 3; c++) { img->pitch[c] = w[c]; bytes[c] = size_t(w[c]) * h[c]; }
            return 3;
        case ROCJPEG_OUTPUT_NATIVE:
            switch (css) {
                case ROCJPEG_CSS_444: case ROCJPEG_CSS_440: for (int c = 0; c < 3; c++) { img->pitch[c] = w[c]; bytes[c] = size_t(w[c]) * h[c]; } return 3;
                case ROCJPEG_CSS_422: img->pitch[0] = 2 * W; bytes[0] = size_t(2) * W * H; return 1;
                case ROCJPEG_CSS_420: img->pitch[0] = W; bytes[0] = size_t(W) * H; img->pitch[1] = W; bytes[1] = size_t(W) * (H >> 1); return 2;
                case ROCJPEG_CSS_400: img->pitch[0] = W; bytes[0] = size_t(W) * H; return 1;
                default: return 0;
            }
        default: return 0;
    }
}

int main(int argc, char** argv) {
    std::string input, fmt_name = "rgb";
    int batch = 32, io_threads = 8, device = 0, passes = 1;
    for (int i = 1; i < argc; i++) {
        auto arg = [&](const char* name) { return !std::strcmp(argv[i], name) && i + 1 < argc; };
        if (arg("-i")) input = argv[++i];
        else if (arg("-fmt")) fmt_name = argv[++i];
        else if (arg("-b")) batch = std::max(1, std::atoi(argv[++i]));
        else if (arg("-t")) io_threads = std::atoi(argv[++i]);
        else if (arg("-d")) device = std::atoi(argv[++i]);
        else if (arg("-n")) passes = std::max(1, std::atoi(argv[++i]));
        else { std::fprintf(stderr, "usage: %s -i <dir|file> [-fmt native|yuv_planar|y|rgb|rgb_planar] [-b batch] [-t io_threads] [-d device] [-n passes]\n", argv[0]); return 2; }
    }
    RocJpegDecodeParams params = {};
    if (fmt_name == "native") params.output_format = ROCJPEG_OUTPUT_NATIVE;
    else if (fmt_name == "yuv_planar") params.output_format = ROCJPEG_OUTPUT_YUV_PLANAR;
    else if (fmt_name == "y") params.output_format = ROCJPEG_OUTPUT_Y;
    else if (fmt_name == "rgb") params.output_format = ROCJPEG_OUTPUT_RGB;
    else if (fmt_name == "rgb_planar") params.output_format = ROCJPEG_OUTPUT_RGB_PLANAR;
    else { std::fprintf(stderr, "unknown output format %s\n", fmt_name.c_str()); return 2; }
    std::vector<std::string> files;
    if (std::filesystem::is_directory(input)) {
        for (const auto& e : std::filesystem::directory_iterator(input))
            if (e.is_regular_file()) files.push_back(e.path().string());
        std::sort(files.begin(), files.end());
    } else if (!input.empty()) {
        files.push_back(input);
    }
    if (files.empty()) { std::fprintf(stderr, "no input files\n"); return 2; }
    if (cudaSetDevice(device) != cudaSuccess) { std::fprintf(stderr, "cannot use device %d\n", device); return 1; }
    RocJpegHandle handle = nullptr;
    CHECK(rocJpegCreate(ROCJPEG_BACKEND_HARDWARE, device, &handle));
    std::vector<RocJpegStreamHandle> streams(static_cast<size_t>(batch));
    for (auto& s : streams) CHECK(rocJpegStreamCreate(&s));
    std::vector<RocJpegImage> outputs(static_cast<size_t>(batch));
    std::vector<size_t> capacity(static_cast<size_t>(batch) * 3, 0);
    std::vector<RocJpegStatus> per_file(static_cast<size_t>(batch));
    std::vector<const char*> paths(static_cast<size_t>(batch));
    size_t decoded = 0, skipped = 0;
    double pixels = 0, seconds = 0;
    for (int pass = 0; pass < passes; pass++) {
        for (size_t first = 0; first < files.size(); first += size_t(batch)) {
            const int n = int(std::min(files.size() - first, size_t(batch)));
            for (int i = 0; i < n; i++) paths[size_t(i)] = files[first + size_t(i)].c_str();
            const auto t0 = std::chrono::steady_clock::now();
            rocJpegB200StreamLoadFiles(streams.data(), paths.data(), n, io_threads, per_file.data());
            // keep the pictures this library decodes (the reference's samples skip unsupported ones the same way)
            std::vector<RocJpegStreamHandle> good;
            std::vector<RocJpegImage> dst;
            double px = 0;
            for (int i = 0; i < n; i++) {
                uint8_t ncomp = 0;
                RocJpegChromaSubsampling css;
                uint32_t w[ROCJPEG_MAX_COMPONENT] = {}, h[ROCJPEG_MAX_COMPONENT] = {};
                RocJpegB200StreamInfo info;
                if (per_file[size_t(i)] != ROCJPEG_STATUS_SUCCESS || rocJpegGetImageInfo(handle, streams[size_t(i)], &ncomp, &css, w, h) != ROCJPEG_STATUS_SUCCESS ||
                    rocJpegB200StreamGetInfo(streams[size_t(i)], &info) != ROCJPEG_STATUS_SUCCESS || info.decode_status != ROCJPEG_STATUS_SUCCESS) {
                    skipped++;
                    continue;
                }
                size_t bytes[3];
                RocJpegImage img;
                if (ChannelSizes(params.output_format, css, w, h, &img, bytes) == 0) { skipped++; continue; }
                RocJpegImage& slot = outputs[size_t(i)];
                for (int c = 0; c < 3; c++) {
                    if (bytes[c] > capacity[size_t(i) * 3 + size_t(c)]) {   // grow-only, as the reference's samples reuse their buffers
                        if (slot.channel[c]) cudaFree(slot.channel[c]);
                        if (cudaMalloc(reinterpret_cast<void**>(&slot.channel[c]), bytes[c]) != cudaSuccess) { std::fprintf(stderr, "out of device memory\n"); return 1; }
                        capacity[size_t(i) * 3 + size_t(c)] = bytes[c];
                    }
                    img.channel[c] = bytes[c] ? slot.channel[c] : nullptr;
                }
                good.push_back(streams[size_t(i)]);
                dst.push_back(img);
                px += double(w[0]) * h[0];
            }
            if (!good.empty()) CHECK(rocJpegDecodeBatched(handle, good.data(), int(good.size()), &params, dst.data()));
            seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            decoded += good.size();
            pixels += px;
        }
    }
    std::printf("Total decoded images: %zu\n", decoded);
    if (skipped) std::printf("Skipped (unreadable / unsupported) files: %zu\n", skipped);
    if (seconds > 0) {
        std::printf("Average processing time per image incl. file read and parse (ms): %.4f\n", 1e3 * seconds / std::max<size_t>(decoded, 1));
        std::printf("Average decoded images per sec (Images/Sec): %.1f\n", double(decoded) / seconds);
        std::printf("Average decoded images size (Mpixels/Sec): %.1f\n", pixels / 1e6 / seconds);
    }
    for (auto& o : outputs)
        for (int c = 0; c < 3; c++)
            if (o.channel[c]) cudaFree(o.channel[c]);
    for (auto& s : streams) rocJpegStreamDestroy(s);
    rocJpegDestroy(handle);
    return 0;
}
