// TEST INFRASTRUCTURE ONLY -- not part of the shipped product path.
//
// A ctypes-friendly C ABI over the *unmodified* reference C++ objects (CPU mode).
// It is compiled together with the reference's own sources, where they lie under
// /root/reference/ModelOptimizations/DlQuantization/src, by oracle/Makefile into
// oracle/_ref/libaimet_ref.so. Nothing from the reference is copied into this repo:
// this file only includes the reference's public headers and forwards calls.
//
// Used by: tests/ (to pin the C restatement in oracle/qsim_oracle.c), the golden-vector
// generator tests/golden/make_golden.py, and bench.py's cpu_baseline / --impl reference leg.
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <random>
#include <tuple>
#include <vector>

#include "DlQuantization/IQuantizationEncodingAnalyzer.hpp"
#include "DlQuantization/ITensorQuantizationSim.h"
#include "DlQuantization/Quantization.hpp"
#include "DlQuantization/QuantizerFactory.hpp"
#include "DlQuantization/TensorQuantizer.h"
#include "math_functions.hpp"

using namespace DlQuantization;

namespace
{
struct Analyzer
{
    std::unique_ptr<IQuantizationEncodingAnalyzer<float>> impl;
};

void put(const TfEncoding& e, double* out5)
{
    out5[0] = e.min;
    out5[1] = e.max;
    out5[2] = e.delta;
    out5[3] = e.offset;
    out5[4] = e.bw;
}
}   // namespace

extern "C"
{
// ---- encoding analyzers (TfEncodingAnalyzer / TfEnhancedEncodingAnalyzer via the reference factory) ----
void* ref_analyzer_new(int quant_mode)
{
    auto* a = new Analyzer;
    a->impl = getEncodingAnalyzerInstance<float>(static_cast<QuantizationMode>(quant_mode));
    return a;
}

void ref_analyzer_free(void* h)
{
    delete static_cast<Analyzer*>(h);
}

void ref_analyzer_update(void* h, const float* data, size_t n)
{
    static_cast<Analyzer*>(h)->impl->updateStats(data, n, COMP_MODE_CPU);
}

void ref_analyzer_compute(void* h, int bw, int sym, int strict, int unsigned_sym, double* out5)
{
    TfEncoding e = static_cast<Analyzer*>(h)->impl->computeEncoding(bw, sym != 0, strict != 0, unsigned_sym != 0);
    put(e, out5);
}

void ref_analyzer_set_percentile(void* h, float percentile)
{
    static_cast<Analyzer*>(h)->impl->setPercentileValue(percentile);
}

// returns the number of buckets written (512, or 0 if the PDF was never initialised)
int ref_analyzer_histogram(void* h, double* x_left, double* pdf)
{
    auto hist = static_cast<Analyzer*>(h)->impl->getStatsHistogram();
    int i     = 0;
    for (auto& t: hist)
    {
        x_left[i] = std::get<0>(t);
        pdf[i]    = std::get<1>(t);
        ++i;
    }
    return i;
}

// ---- TensorQuantizationSim<float> ----
void ref_fill_encoding_info(int bw, double enc_min, double enc_max, double* out5)
{
    auto sim = getTensorQuantizationSim<float>();
    TfEncoding e;
    sim->fillEncodingInfo(e, bw, enc_min, enc_max);
    put(e, out5);
}

void ref_qdq(const float* in, size_t n, float* out, double enc_min, double enc_max, int bw, int round_mode)
{
    auto sim = getTensorQuantizationSim<float>();
    sim->quantizeDequantizeTensor(in, n, out, enc_min, enc_max, bw, static_cast<RoundingMode>(round_mode), false);
}

void ref_quantize(const float* in, size_t n, float* out, double enc_min, double enc_max, int bw, int round_mode,
                  int shift_to_signed)
{
    auto sim = getTensorQuantizationSim<float>();
    sim->quantizeTensor(in, n, out, enc_min, enc_max, bw, static_cast<RoundingMode>(round_mode), false,
                        shift_to_signed != 0);
}

// ITensorQuantizationSim::quantizeTensorPacked (TensorQuantizationSim.cpp:128-139). out: max(bw, 8) / 8 bytes per element.
// Returns the number of bytes written, or -1 where the reference throws (a bitwidth that is not 1, 2, 4, 8, 16 or 32).
int64_t ref_quantize_packed(const float* in, size_t n, uint8_t* out, double enc_min, double enc_max, int bw, int shift_to_signed)
{
    // (the reference throws for other bitwidths from inside its worker threads, i.e. it terminates the process)
    if (!(bw == 1 || bw == 2 || bw == 4 || bw == 8 || bw == 16 || bw == 32))
        return -1;
    auto sim = getTensorQuantizationSim<float>();
    std::vector<uint8_t> packed;
    try
    {
        sim->quantizeTensorPacked(in, n, packed, enc_min, enc_max, bw, ROUND_NEAREST, false, shift_to_signed != 0);
    }
    catch (const std::exception&)
    {
        return -1;
    }
    std::memcpy(out, packed.data(), packed.size());
    return (int64_t) packed.size();
}

void ref_qdq_broadcast(const float* in, float* out, int64_t num_element, int64_t num_dims, const int64_t* input_strides,
                       const int64_t* encoding_strides, const float* enc_min, const float* enc_max,
                       const float* enc_delta, const float* enc_offset)
{
    quantizeDequantizeBroadcast<float>(in, out, num_element, num_dims, input_strides, encoding_strides, enc_min, enc_max,
                                       enc_delta, enc_offset, COMP_MODE_CPU, nullptr);
}

void ref_qdq_per_channel(const float* in, size_t num_channel, size_t num_element, size_t num_element_per_channel,
                         float* out, float* enc_min, float* enc_max, float* enc_delta, float* enc_offset,
                         int round_mode)
{
    auto sim = getTensorQuantizationSim<float>();
    sim->quantizeDequantizeTensorPerChannel(in, num_channel, num_element, num_element_per_channel, out, enc_min,
                                            enc_max, enc_delta, enc_offset, static_cast<RoundingMode>(round_mode),
                                            false);
}

// ---- TensorQuantizer facade ----
void* ref_tq_new(int quant_mode, int round_mode)
{
    return new TensorQuantizer(static_cast<QuantizationMode>(quant_mode), static_cast<RoundingMode>(round_mode));
}

void ref_tq_free(void* h)
{
    delete static_cast<TensorQuantizer*>(h);
}

void ref_tq_set_flags(void* h, int strict, int unsigned_sym)
{
    auto* q = static_cast<TensorQuantizer*>(h);
    q->setStrictSymmetric(strict != 0);
    q->setUnsignedSymmetric(unsigned_sym != 0);
}

void ref_tq_update(void* h, const float* data, size_t n)
{
    static_cast<TensorQuantizer*>(h)->updateStats(data, n, false);
}

void ref_tq_compute(void* h, int bw, int sym, double* out5)
{
    put(static_cast<TensorQuantizer*>(h)->computeEncoding(bw, sym != 0), out5);
}

int ref_tq_is_valid(void* h)
{
    return static_cast<TensorQuantizer*>(h)->isEncodingValid ? 1 : 0;
}

void ref_tq_reset(void* h)
{
    static_cast<TensorQuantizer*>(h)->resetEncodingStats();
}

// returns 0 on success, 1 if the reference threw
int ref_tq_partial(void* h, int bw, double* enc5, int sym, int unsigned_sym, int strict)
{
    TfEncoding e;
    e.min    = enc5[0];
    e.max    = enc5[1];
    e.delta  = enc5[2];
    e.offset = enc5[3];
    e.bw     = static_cast<int>(enc5[4]);
    try
    {
        static_cast<TensorQuantizer*>(h)->computePartialEncoding(bw, e, sym != 0, unsigned_sym != 0, strict != 0);
    }
    catch (const std::exception&)
    {
        return 1;
    }
    put(e, enc5);
    return 0;
}
// ---- TensorQuantizer's per-channel 4-D helpers (TensorQuantizer.cpp:189-326): slice on `axis`, one encoding per slice from
//      its data, quantize-dequantize / pack every slice, concat. enc_out: [shape4[axis]][5]. Return -1 if the reference threw.
int ref_tq_qdq_per_channel_tensor(int quant_mode, const float* in, const uint32_t* shape4, uint32_t axis, int bw, int strict,
                                  float* out, double* enc_out)
{
    try
    {
        TensorQuantizer tq(static_cast<QuantizationMode>(quant_mode), ROUND_NEAREST);
        std::vector<uint32_t> shape(shape4, shape4 + 4);
        std::vector<TfEncoding> encodings;
        tq.quantizeDequantizePerChannelTensor(in, shape, axis, out, encodings, (uint8_t) bw, ROUND_NEAREST, false, strict != 0);
        for (size_t i = 0; i < encodings.size(); ++i)
            put(encodings[i], enc_out + 5 * i);
        return (int) encodings.size();
    }
    catch (const std::exception&)
    {
        return -1;
    }
}

int64_t ref_tq_packed_per_channel_tensor(int quant_mode, const float* in, const uint32_t* shape4, uint32_t axis, int bw,
                                         int strict, uint8_t* out, double* enc_out)
{
    try
    {
        TensorQuantizer tq(static_cast<QuantizationMode>(quant_mode), ROUND_NEAREST);
        std::vector<uint32_t> shape(shape4, shape4 + 4);
        std::vector<TfEncoding> encodings;
        std::vector<uint8_t> packed;
        tq.quantizePerChannelTensorPacked(in, shape, axis, packed, encodings, (uint8_t) bw, ROUND_NEAREST, false, strict != 0);
        for (size_t i = 0; i < encodings.size(); ++i)
            put(encodings[i], enc_out + 5 * i);
        std::copy(packed.begin(), packed.end(), out);
        return (int64_t) packed.size();
    }
    catch (const std::exception&)
    {
        return -1;
    }
}

// The input the reference's own fixture builds (DlQuantization/test/TestTensorQuantizer.cpp:92-103):
// std::normal_distribution<float>(mean, stddev) driven by std::mt19937(seed). libstdc++-specific, hence generated here.
// ---- the entropy scheme's raw histogram: updateTensorHistogram on a TensorProfilingParams the test owns
//      (EntropyEncodingAnalyzer keeps its own private; the analyzer object itself is reachable through ref_analyzer_*) ----
void* ref_tpp_new()
{
    auto* t       = new TensorProfilingParams;
    t->min        = 0;
    t->max        = 0;
    t->iterations = 0;
    return t;
}
void ref_tpp_free(void* h)
{
    delete static_cast<TensorProfilingParams*>(h);
}
void ref_tpp_update(void* h, const float* data, int n)
{
    updateTensorHistogram(data, n, COMP_MODE_CPU, *static_cast<TensorProfilingParams*>(h));
}
// returns the histogram size (0 before the first non-zero batch)
int ref_tpp_get(void* h, double* hist, double* min_max, int* iterations)
{
    auto* t    = static_cast<TensorProfilingParams*>(h);
    min_max[0] = t->min;
    min_max[1] = t->max;
    *iterations = t->iterations;
    for (size_t i = 0; i < t->histogram.size(); ++i)
        hist[i] = t->histogram[i];
    return (int) t->histogram.size();
}
void ref_rescale_histogram(const double* src, int n, double src_min, double src_max, double dst_min, double dst_max,
                           double* dst)
{
    std::vector<double> s(src, src + n);
    std::vector<double> d = rescaleHistogram(s, src_min, src_max, dst_min, dst_max);
    for (int i = 0; i < n; ++i)
        dst[i] = d[i];
}

void ref_kat_normal(unsigned seed, float mean, float stddev, unsigned n, float* out)
{
    std::normal_distribution<float> distribution(mean, stddev);
    std::mt19937 generator(seed);
    for (unsigned i = 0; i < n; i++)
        out[i] = distribution(generator);
}
}   // extern "C"
