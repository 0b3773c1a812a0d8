/*
 * aimet_b200 -- C ABI of the B200-native quantization-simulation hot path.
 *
 * This header is the drop-in boundary: every entry point replaces one call the
 * reference's two Python-visible native modules make into ModelOptimizations/DlQuantization. All
 * file:line citations are into the reference checkout (shayaanjamil-10xe/aimet, v1.35.0):
 *   DlQ = ModelOptimizations/DlQuantization,  ATQ = TrainingExtensions/torch/src/AimetTensorQuantizer.cpp
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary
 *   - `const void* in` / `void* out` are DEVICE pointers to contiguous fp32 (AB_F32) or bf16 (AB_BF16) data
 *   - element counts are int64_t (the reference uses `int cnt`, which overflows at 2^31 elements)
 *   - every device entry point enqueues work on `stream` (a cudaStream_t passed as void*) and returns
 *     without synchronising; nothing allocates device memory: state and scratch are caller-owned
 *   - return value: AB_OK (0) or a negative ab_status; ab_last_error() gives a thread-local message
 *   - there is no CPU fallback: on a machine without a CUDA device every device entry point fails
 *     with AB_ERR_CUDA
 */
#ifndef AIMET_B200_H_
#define AIMET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AB_PDF_SIZE 512 /* DlQ/src/math_functions.hpp:80 (PDF_SIZE) */

typedef enum
{
    AB_OK              = 0,
    AB_ERR_INVALID     = -1, /* bad argument (the reference throws std::invalid_argument / runtime_error) */
    AB_ERR_CUDA        = -2, /* CUDA runtime error, or no device */
    AB_ERR_UNSUPPORTED = -3
} ab_status;

typedef enum
{
    AB_F32  = 0,
    AB_BF16 = 1
} ab_dtype;

/* DlQ/include/DlQuantization/Quantization.hpp:76-80 */
typedef enum
{
    AB_ROUND_NEAREST    = 0,
    AB_ROUND_STOCHASTIC = 1
} ab_round_mode;

/* DlQ/include/DlQuantization/Quantization.hpp:83-107 (the schemes on the hot path, with the reference's values) */
typedef enum
{
    AB_QUANTIZATION_TF          = 0,
    AB_QUANTIZATION_TF_ENHANCED = 1,
    AB_QUANTIZATION_PERCENTILE  = 3, /* statistics identical to TF_ENHANCED (UpdatePdf); see ab_compute_encodings_percentile */
    AB_QUANTIZATION_MSE         = 4  /* MseEncodingAnalyzer (DlQ/src/MseEncodingAnalyzer.cpp): same statistics again */
} ab_quant_mode;

/* DlQ/include/DlQuantization/Quantization.hpp:113-120 (TfEncoding), same field order */
typedef struct
{
    double min;
    double max;
    double delta;
    double offset;
    int bw;
} ab_encoding;

/* Per-quantizer statistics record, resident in DEVICE memory (one per tensor quantizer, or one per channel).
 * It replaces the host-side state of TfEncodingAnalyzer (DlQ/src/TfEncodingAnalyzer.h:85-93: running min/max)
 * and TfEnhancedEncodingAnalyzer (DlQ/src/math_functions.hpp:59-67: PDF{xLeft, pdf, iterations}).
 * xLeft[i] is not stored: it is x_left0 + i * bucket_size_d, exactly as InitializePdf computes it
 * (DlQ/src/math_functions.cpp:231-236). Plain old data, sizeof is a multiple of 16; zero bytes are NOT a valid initial
 * state: use ab_stats_reset().
 */
typedef struct
{
    double pdf[AB_PDF_SIZE];       /* running mean of per-batch PDFs (math_functions.cpp:279-287) */
    uint32_t hist[2][AB_PDF_SIZE]; /* raw counts of the latest batch(es); see `pending` */
    double x_left0;                /* xLeft[0] */
    double bucket_size_d;          /* (max - min) / 512 in double (math_functions.cpp:222) */
    double run_min;                /* TF scheme: running min over batches (TfEncodingAnalyzer.cpp:69) */
    double run_max;                /* TF scheme: running max over batches (TfEncodingAnalyzer.cpp:70) */
    double pending_count;          /* element count of the batch whose counts still await their fold */
    float bucket_size;             /* float(xLeft[1] - xLeft[0]) (math_functions.cpp:265) */
    float pdf_offset;              /* float(xLeft[0]) / bucket_size (math_functions.cpp:266-268) */
    int32_t batch_min_bits;        /* scratch: order-preserving int image of the current batch's min */
    int32_t batch_max_bits;        /* scratch: ... max */
    int32_t initialized;           /* PDF range fixed (xLeft.size() != 0) */
    int32_t stats_updated;         /* updateStats was called at least once (_statsUpdated) */
    int32_t iterations;            /* PDF.iterations (batches already folded into pdf) */
    uint32_t ticket;               /* scratch: last-block election counter; zero between calls */
    int32_t pending;               /* 1: hist[write_parity ^ 1] holds a batch that is NOT yet folded into pdf. The
                                      histogram kernel leaves the fold (pdf = (pdf*k + hist/cnt)/(k+1)) of batch k to the
                                      next call on this record, where it overlaps the streaming of batch k+1 instead of
                                      sitting on the kernel's tail; every reader folds it on the fly. */
    int32_t write_parity;          /* which hist[] buffer the next batch's counts go to */
    /* bf16 tensors only: a certified one-FMA form of the bin index. Once the range is frozen, a statistics call on a
     * large bf16 tensor compares  floor(fma(x, scale, shift))  with the reference's  round(x / bucket - offset)  for
     * ALL 65536 bf16 bit patterns (a few per warp, off the critical path), for nine (scale, shift) candidates around
     * (1 / bucket, 0.5 - offset); later calls use the first candidate that reproduced every pattern (6 instructions
     * per sample instead of 14). If none did, the exact sequence stays. */
    float bf16_scale;
    float bf16_shift;
    int32_t bf16_formula;          /* 0: not examined yet, 1: (bf16_scale, bf16_shift) certified, -1: no candidate is exact */
    uint32_t bf16_fail_mask;       /* scratch while certifying: bit j = candidate j misplaced some pattern; zero between calls */
} ab_stats_state;

const char* ab_last_error(void);
/* library / device probes (host only) */
int ab_version(void);
int ab_device_count(void);
size_t ab_stats_state_bytes(void); /* == sizeof(ab_stats_state) */
/* Measurement hook, not used in normal operation. slots: DEVICE array of capacity x 3 uint64, each triple initialised to
 * {INT64_MAX, 0, 0}. Every histogram launch (tf_enhanced ab_stats_update) issued afterwards takes the next triple and
 * records the earliest start / latest end of its CTAs on the GPU's global nanosecond timer and the bytes of its input:
 * the launch's execution time on the device, free of launch latency and of the event records a host-side bracket needs.
 * slots == NULL switches it off. Returns the number of triples handed out since the previous call. Not thread-safe. */
int64_t ab_debug_hist_timer(unsigned long long* slots, int64_t capacity);

/* ------------------------------------------------------------------------------------------------------------
 * Host helpers: encoding math in double, bit-identical to the reference's host code.
 * ---------------------------------------------------------------------------------------------------------- */

/* gateMinMax: DlQ/src/quantization_utils.cpp:145-156 */
int ab_gate_min_max(double* enc_min, double* enc_max);
/* TensorQuantizationSim::fillEncodingInfo / generateScaleOffset: DlQ/src/TensorQuantizationSim.cpp:63-92 */
int ab_fill_encoding_info(int bw, double enc_min, double enc_max, ab_encoding* out);
/* getComputedEncodings (TF scheme, from a min/max pair): DlQ/src/quantization_utils.cpp:58-143 */
int ab_tf_compute_encoding(int bw, double mn, double mx, int use_symmetric, int use_strict_symmetric,
                           int use_unsigned_symmetric, ab_encoding* out);
/* TfEncodingAnalyzer::computeEncoding (adds the include-zero / MIN_RANGE gating): DlQ/src/TfEncodingAnalyzer.cpp:81-101 */
int ab_tf_analyzer_encoding(int bw, double run_min, double run_max, int use_symmetric, int use_strict_symmetric,
                            int use_unsigned_symmetric, ab_encoding* out);
/* TensorQuantizer::computePartialEncoding: DlQ/src/TensorQuantizer.cpp:327-343 -> quantization_utils.cpp:158-228.
 * AB_ERR_INVALID where the reference throws. */
int ab_compute_partial_encoding(int bw, ab_encoding* enc, int use_symmetric, int use_unsigned_symmetric,
                                int use_strict_symmetric);
/* The per-channel parameter preparation AimetTensorQuantizer::quantizeDequantizePerChannel does with torch fp32
 * CPU ops (ATQ:236-299): gate, delta = (max-min)/steps, offset = rint(min/delta); the step count is decided from
 * channel 0 only. Writes params[0..C) = min, [C..2C) = max, [2C..3C) = delta, [3C..4C) = offset (host memory). */
int ab_per_channel_params(const double* enc_min, const double* enc_max, int num_channel, int bw, float* params);

/* ------------------------------------------------------------------------------------------------------------
 * Job 1: fused quantize-dequantize, quantize-only, and straight-through-estimator backward.
 * ---------------------------------------------------------------------------------------------------------- */

/* ITensorQuantizationSim::quantizeDequantizeTensor (DlQ/src/TensorQuantizationSim.cpp:104-114 ->
 * trim_functions.cpp:94-113,174-182 ; GPU twin trim_functions.cu:46-60,126-132).
 * y = delta * (round(clamp(x, min, max) / delta - offset) + offset) with the encoding derived from
 * (enc_min, enc_max, bw) by fillEncodingInfo. bf16: x is widened to fp32, y is rounded to bf16 (RNE), which is
 * what the Python host does around the call (aimet_torch/v1/tensor_quantizer.py:1129-1136). */
int ab_qdq_per_tensor_fwd(const void* in, void* out, int64_t count, int dtype, double enc_min, double enc_max,
                          int bw, int round_mode, uint64_t seed, void* stream);

/* Same kernel, but the four fp32 parameters {min, max, delta, offset} are read from DEVICE memory (`enc4`),
 * so an encoding produced on the device (ab_compute_encodings) can be consumed without a host round trip. */
int ab_qdq_per_tensor_fwd_dev(const void* in, void* out, int64_t count, int dtype, const float* enc4,
                              int round_mode, uint64_t seed, void* stream);

/* ITensorQuantizationSim::quantizeTensor (DlQ/src/TensorQuantizationSim.cpp:116-126 -> trim_functions.cpp:202-218;
 * GPU twin trim_functions.cu:62-76,151-162): the integer grid value, stored in the float type, minus
 * 2^(bw-1) when shift_to_signed. */
int ab_quantize_to_grid(const void* in, void* out, int64_t count, int dtype, double enc_min, double enc_max, int bw,
                        int round_mode, int shift_to_signed, uint64_t seed, void* stream);

/* ITensorQuantizationSim::quantizeTensorPacked (DlQ/src/TensorQuantizationSim.cpp:128-139 -> quantizeToFxpPackedCpu,
 * DlQ/src/trim_functions.cpp:221-388; "GPU packed quantization not supported" in the reference, :194-196): the integer grid
 * value of every element, nearest rounding, computed in double on the double encoding, stored in `out` (DEVICE) as
 * max(bw, 8) / 8 bytes per element: uint8 (one value per byte below 8 bit, as the reference stores them), uint16 or uint32;
 * with shift_to_signed the value minus 2^(bw-1) - 1 as int8 (low bw bits below 8 bit) / int16 / int32.
 * bw must be 1, 2, 4, 8, 16 or 32 (AB_ERR_INVALID otherwise: the reference throws). */
int ab_quantize_to_packed(const void* in, void* out, int64_t count, int dtype, double enc_min, double enc_max, int bw,
                          int shift_to_signed, void* stream);

/* ITensorQuantizationSim::quantizeDequantizeTensorPerChannel (DlQ/src/TensorQuantizationSim.cpp:281-318 ->
 * trim_functions.cpp:697-709 ; GPU twin trim_functions.cu:78-92,169-172).
 * channel(i) = (i / num_element_per_channel) % num_channel. `params` is a DEVICE array of 4*num_channel floats
 * laid out as ab_per_channel_params writes it. */
int ab_qdq_per_channel_fwd(const void* in, void* out, int64_t num_channel, int64_t num_element,
                           int64_t num_element_per_channel, int dtype, const float* params, int round_mode,
                           uint64_t seed, void* stream);

/* ab_per_channel_params on the device: `enc5` is a DEVICE array of num_channel x 5 doubles {min, max, delta, offset, bw} as
 * ab_compute_encodings writes it; `params` a DEVICE array of 4*num_channel floats. Lets freshly searched per-channel
 * encodings feed ab_qdq_per_channel_fwd without a host round trip (same arithmetic: ATQ:236-299). */
int ab_per_channel_params_dev(const double* enc5, int64_t num_channel, int bw, float* params, void* stream);

/* quantizeDequantizeBroadcast (DlQ/include/DlQuantization/Quantization.hpp:193-221 -> DlQ/src/trim_functions.cpp:633-662; GPU
 * twin trim_functions.cu:96-122; caller: the ONNX custom op, TrainingExtensions/onnx/src/AimetOpUtils.h:269): QDQ with an
 * encoding tensor broadcast over the input (blockwise / LPBQ, several per-channel axes ...). The input is a contiguous
 * tensor of `num_dims` (<= 8) dimensions whose element strides are `input_strides`; `encoding_strides` are the element
 * strides of the encoding tensor padded to the same rank, 0 along broadcast dimensions. Both stride arrays are HOST arrays
 * here (they travel as kernel arguments; the reference wants them in device memory). enc_min / enc_max / enc_delta /
 * enc_offset: DEVICE float arrays holding the encoding tensor, used as they are (no gating), rounding to nearest. */
int ab_qdq_broadcast_fwd(const void* in, void* out, int64_t num_element, int num_dims, const int64_t* input_strides,
                         const int64_t* encoding_strides, const float* enc_min, const float* enc_max,
                         const float* enc_delta, const float* enc_offset, int dtype, void* stream);

/* compute_dloss_by_dx (TrainingExtensions/torch/src/python/aimet_torch/v1/quantsim_straight_through_grad.py:91-118):
 * grad_in = grad * [enc_min <= x <= enc_max]. x, grad, grad_in share `dtype`. */
int ab_qdq_ste_bwd(const void* x, const void* grad, void* grad_in, int64_t count, int dtype, float enc_min,
                   float enc_max, void* stream);
/* per-channel variant: enc_min / enc_max are DEVICE arrays of num_channel floats */
int ab_qdq_ste_bwd_per_channel(const void* x, const void* grad, void* grad_in, int64_t num_channel,
                               int64_t num_element, int64_t num_element_per_channel, int dtype,
                               const float* enc_min, const float* enc_max, void* stream);

/* same, with the range taken from the rows {min, max, delta, offset, bw} (num_channel x 5 doubles, DEVICE) that
 * ab_compute_encodings leaves behind: float32(min) <= x <= float32(max), the narrowing `torch.tensor(python float)` does
 * in the reference; range_in_bf16 != 0 rounds the two bounds once more to bf16 (the reference compares a bf16 tensor with a
 * 0-dim bound in the tensor's dtype). Saves the host the slicing / casting kernels in front of every backward. */
int ab_qdq_ste_bwd_enc5(const void* x, const void* grad, void* grad_in, int64_t num_channel, int64_t num_element,
                        int64_t num_element_per_channel, int dtype, const double* enc5, int range_in_bf16, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Job 2: statistics -- min/max and the 512-bin histogram, accumulated into device-resident state.
 * ---------------------------------------------------------------------------------------------------------- */

/* (re)initialise `count` records: IQuantizationEncodingAnalyzer construction / resetEncodingStats (ATQ:85-92) */
int ab_stats_reset(ab_stats_state* states, int64_t count, void* stream);

/* One updateStats call on one quantizer (ATQ:94-126).
 *  AB_QUANTIZATION_TF          : TfEncodingAnalyzer::updateStats (DlQ/src/TfEncodingAnalyzer.cpp:60-71): one pass,
 *                                running min / max.
 *  AB_QUANTIZATION_TF_ENHANCED : UpdatePdf (DlQ/src/math_functions.cpp:243-288): on the first non-zero batch a min/max
 *                                pass fixes the histogram range (InitializePdf :207-241), then GetHistogram (:367-384)
 *                                bins the batch and the counts are folded into the running PDF -- all on the device,
 *                                with no host synchronisation (two launches; one once the range is fixed).
 * `batch_log_entry`, if not NULL, is a DEVICE array of AB_PDF_SIZE + 2 uint32, ZEROED BY THE CALLER, to which this batch's
 * raw counts are added, followed by the element count (low word, high word; 0 when the batch was skipped because the PDF
 * was still uninitialised and the batch was all zeros). Used by the multi-GPU exact merge (ab_stats_fold_batches).
 * `flags`: AB_STATS_RANGE_FIXED -- the caller KNOWS (from an earlier read-back of `initialized`) that this record's
 * histogram range is fixed, so the min/max kernel, which would exit immediately, is not even launched. */
#define AB_STATS_RANGE_FIXED 1
int ab_stats_update(const void* in, int64_t count, int dtype, int quant_mode, ab_stats_state* state,
                    uint32_t* batch_log_entry, int flags, void* stream);

/* updateStats on `num_segments` quantizers at once: segment s is in[s*segment_len .. (s+1)*segment_len) and updates
 * states[s]. Replaces the per-channel Python loop (aimet_torch/v1/tensor_quantizer.py:567-570). */
int ab_stats_update_segmented(const void* in, int64_t num_segments, int64_t segment_len, int dtype, int quant_mode,
                              ab_stats_state* states, void* stream);

/* Many tf_enhanced updateStats calls -- each on its own tensor and its own record -- as ONE histogram launch plus one tiny
 * fold launch (net-new: the reference makes one native call, i.e. several kernels and a blocking copy, per tensor:
 * ATQ:94-126 -> DlQ/src/math_functions.cu:125-211). Every record addressed must have its histogram range fixed
 * (`initialized`, i.e. a call on a record without a range is not counted); the host layer sends first batches through
 * ab_stats_update. Records may repeat (a module called several times per forward): their calls are folded in table order.
 *   segments   : HOST array (the table travels as a kernel argument, so the call is capturable in a CUDA graph)
 *   dtype      : element type of ALL segments of this call
 *   states     : DEVICE base of the records; segment s updates states[segments[s].state_index]
 *   seg_counts : DEVICE uint32 [num_segments][AB_PDF_SIZE + 2], ZEROED by the caller. Without flags the rows are scratch:
 *                the fold consumes them and hands them back zeroed. With AB_STATS_MULTI_LOG_ONLY nothing is folded: row s
 *                keeps the raw counts of call s followed by its element count (low, high word) -- the log entry format of
 *                ab_stats_update / ab_stats_fold_batches for the multi-GPU exact merge. */
typedef struct
{
    const void* data;    /* DEVICE, contiguous, 16-byte aligned */
    int64_t count;       /* elements (at least one 128-bit vector) */
    int32_t state_index; /* record index relative to `states` */
    int32_t reserved;
} ab_stats_segment;
#define AB_STATS_MULTI_MAX_SEGMENTS 128
#define AB_STATS_MULTI_LOG_ONLY 1
#define AB_STATS_MULTI_HIST_ONLY 2 /* measurement: enqueue only the histogram launch ... */
#define AB_STATS_MULTI_FOLD_ONLY 4 /* ... and only the fold launch (same table), so that a caller can bracket either */
int ab_stats_update_multi(const ab_stats_segment* segments, int num_segments, int dtype, ab_stats_state* states,
                          uint32_t* seg_counts, int flags, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Job 3: tf_enhanced grid search on the device.
 * ---------------------------------------------------------------------------------------------------------- */

/* TfEnhancedEncodingAnalyzer::computeEncoding for `count` quantizers in one launch
 * (DlQ/src/TfEnhancedEncodingAnalyzer.cpp:79-113, 115-144, 178-253, 256-291, 294-355, 358-397).
 * enc_out   : DEVICE array of count * 5 doubles {min, max, delta, offset, bw}; all zero when no stats were seen
 * qdq4_out  : optional DEVICE array of count * 4 floats: the {min, max, delta, offset} fp32 kernel parameters that
 *             fillEncodingInfo(enc.min, enc.max, bw) yields (per-tensor QDQ, for ab_qdq_per_tensor_fwd_dev); may be NULL
 * AB_QUANTIZATION_TF states are also accepted (then the encoding is TfEncodingAnalyzer::computeEncoding), and so is
 * AB_QUANTIZATION_MSE (MseEncodingAnalyzer<float>::computeEncoding, DlQ/src/MseEncodingAnalyzer.cpp:77-285, on
 * tf_enhanced-style statistics). */
int ab_compute_encodings(const ab_stats_state* states, int64_t count, int quant_mode, int bw, int use_symmetric,
                         int use_strict_symmetric, int use_unsigned_symmetric, double* enc_out, float* qdq4_out,
                         void* stream);

/* One tensor's resetEncodingStats + updateStats + computeEncoding enqueued by a single call: what a parameter quantizer does
 * before every forward in training mode (TrainingExtensions/torch/src/python/aimet_torch/v1/qc_quantize_op.py:753-798).
 * num_segments == 1: the tensor is one quantizer's input of segment_len elements (ab_stats_update); otherwise segment s
 * updates states[s] (ab_stats_update_segmented). enc_out / qdq4_out as in ab_compute_encodings; params_out (optional):
 * the float[4][num_segments] block of ab_per_channel_params_dev. Same kernels, same results as the separate calls. */
int ab_stats_refresh_encodings(const void* in, int64_t num_segments, int64_t segment_len, int dtype, int quant_mode,
                               ab_stats_state* states, int bw, int use_symmetric, int use_strict_symmetric,
                               int use_unsigned_symmetric, double* enc_out, float* qdq4_out, float* params_out,
                               void* stream);

/* The same for MANY tensors with a handful of launches in total (one reset, one statistics launch over all segments of all
 * tensors, one grid search over all records, one parameter-block launch) instead of four per tensor: the parameter
 * quantizers of a whole model before a training-mode forward / once per calibration job. All tensors share `dtype`, the
 * scheme and the encoding flags. Item i's records are states[first_record .. first_record + num_segments); the items must
 * tile the record range [0, total) in order. enc_out: total x 5 doubles; qdq4_out (optional): total x 4 floats;
 * params_out (optional): 4 * total floats, item i's float[4][num_segments] block at params_out + 4 * first_record.
 * Same kernels' arithmetic, same results as ab_stats_refresh_encodings per tensor. */
typedef struct
{
    const void* data;     /* DEVICE, contiguous: num_segments x segment_len elements */
    int64_t num_segments; /* 1: one per-tensor quantizer; C: one record per channel */
    int64_t segment_len;
    int64_t first_record;
} ab_refresh_item;
#define AB_REFRESH_MULTI_MAX_ITEMS 96
int ab_stats_refresh_encodings_multi(const ab_refresh_item* items, int num_items, int dtype, int quant_mode,
                                     ab_stats_state* states, int bw, int use_symmetric, int use_strict_symmetric,
                                     int use_unsigned_symmetric, double* enc_out, float* qdq4_out, float* params_out,
                                     void* stream);

/* PercentileEncodingAnalyzer<float>::computeEncoding (DlQ/src/PercentileEncodingAnalyzer.cpp:77-196) for `count` consecutive
 * records whose statistics were collected with AB_QUANTIZATION_PERCENTILE (or TF_ENHANCED: the PDF is the same).
 * `percentile` is what setPercentileValue received (:203-206; 100 = the observed range). Outputs as ab_compute_encodings. */
int ab_compute_encodings_percentile(const ab_stats_state* states, int64_t count, float percentile, int bw,
                                    int use_symmetric, int use_strict_symmetric, int use_unsigned_symmetric,
                                    double* enc_out, float* qdq4_out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Range learning ("learned grid" QAT): fused forward and backward of QuantizeDequantizeFunc
 * (TrainingExtensions/torch/src/python/aimet_torch/v1/tensor_quantizer.py:854-963), i.e. of
 * get_computed_encodings / calculate_forward_pass / asymmetric_gradients / symmetric_gradients
 * (v1/quantsim_straight_through_grad.py:121-346) and, with AB_LG_GATE, of set_encoding_min_max_gating_threshold
 * (v1/tensor_quantizer.py:1347-1359).
 *
 * The tensor is viewed as [outer][num_channel][inner] (per-tensor: num_channel == 1), channel(i) = (i / inner) %
 * num_channel, which is broadcast_to_tensor (quantsim_straight_through_grad.py:70-92) for any channel axis.
 * enc_min / enc_max (and grad_min / grad_max) are DEVICE arrays of num_channel elements of the tensor's own `dtype`
 * (the reference requires tensor and encoding parameters to share a dtype, :199-202). Arithmetic follows the
 * reference's dtype rule: fp32 tensors in fp32; bf16 tensors in bf16 (every operation rounded) below 16 bit and in fp32
 * from 16 bit up (:211-214).
 * ---------------------------------------------------------------------------------------------------------- */
enum ab_lg_symmetry
{
    AB_LG_ASYMMETRIC         = 0,
    AB_LG_SIGNED_SYMMETRIC   = 1, /* use_symmetric_encodings and not is_unsigned_symmetric */
    AB_LG_UNSIGNED_SYMMETRIC = 2
};
#define AB_LG_GATE 1 /* flags: clamp (min, max) in place first, as LearnedGridQuantWrapper.apply_gating_logic does */

/* bytes of device scratch ab_lg_qdq_fwd (per-channel only) and ab_lg_qdq_bwd need for `num_channel` channels. The
 * caller zero-fills it ONCE; the kernels leave it zeroed again. One workspace per stream. */
int64_t ab_lg_workspace_bytes(int64_t num_channel);

/* y = (clamp(round_half_even(x / delta) - offset, 0, num_steps) + offset) * delta with (delta, offset, num_steps) derived
 * on the device from (enc_min, enc_max). `workspace` may be NULL when num_channel == 1. */
int ab_lg_qdq_fwd(const void* in, void* out, int64_t outer, int64_t num_channel, int64_t inner, int dtype,
                  void* enc_min, void* enc_max, int bw, int sym_mode, int use_strict_symmetric, int flags,
                  void* workspace, void* stream);

/* grad_in = grad * mask (skipped when grad_in is NULL) and, unless grad_min / grad_max are NULL, the gradients of the
 * loss with respect to enc_min / enc_max. `in` is the forward's input, enc_min / enc_max the values the forward used. */
int ab_lg_qdq_bwd(const void* in, const void* grad, void* grad_in, int64_t outer, int64_t num_channel, int64_t inner,
                  int dtype, const void* enc_min, const void* enc_max, int bw, int sym_mode, int use_strict_symmetric,
                  void* grad_min, void* grad_max, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Multi-GPU exact merge (net-new). The reference's tf_enhanced result depends on the range
 * fixed by the first non-zero batch and on a sequential running mean over batches, so ranks exchange integer
 * per-batch histograms and replay them in global batch order.
 * ---------------------------------------------------------------------------------------------------------- */

/* Fix the histogram range of `count` quantizers from given batch (min, max) pairs -- InitializePdf
 * (DlQ/src/math_functions.cpp:207-241) on the device. minmax: DEVICE [count][2] floats; a (0, 0) pair, or a state
 * whose range is already fixed, is left untouched (math_functions.cpp:254-259). */
int ab_stats_init_range(ab_stats_state* states, int64_t count, const float* minmax, void* stream);

/* Rebuild the running PDFs of `count` quantizers from logged batches, replaying
 * pdf = (pdf*k + hist/cnt)/(k+1) (DlQ/src/math_functions.cpp:279-287) for k = 0 .. num_batches-1 in order.
 * batch_log    : DEVICE uint32 buffer holding, for every batch, `count` consecutive entries of AB_PDF_SIZE + 2 words
 *                as ab_stats_update writes them
 * batch_offsets: DEVICE int64[num_batches], the word offset of batch k's first entry inside batch_log
 * Entries whose element count is 0 are skipped, as the reference skips all-zero batches seen before the range is fixed. */
int ab_stats_fold_batches(ab_stats_state* states, int64_t count, const uint32_t* batch_log,
                          const int64_t* batch_offsets, int64_t num_batches, void* stream);

/* The same replay over a log of per-CALL entries (what ab_stats_update with a log entry and ab_stats_update_multi with
 * AB_STATS_MULTI_LOG_ONLY write: AB_PDF_SIZE counts + the element count), each record replaying its own entries in the
 * order given:
 * log          : DEVICE uint32 buffer of rows of AB_PDF_SIZE + 2 words (e.g. the all-gathered logs of all ranks)
 * entry_rows   : DEVICE int64[], row numbers inside `log`, grouped by record, each group in replay order
 * record_begin : DEVICE int64[count + 1], CSR index: record s replays entry_rows[record_begin[s] .. record_begin[s+1]) */
int ab_stats_fold_log(ab_stats_state* states, int64_t count, const uint32_t* log, const int64_t* entry_rows,
                      const int64_t* record_begin, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * The compute body of the reference's ONNX Runtime custom op "QcQuantizeOp" (TrainingExtensions/onnx/src/QcQuantizeOp.cpp:
 * 64-143 -> AimetOpUtils.h:98-322): what QcQuantizeOpCuda::Compute does between its input and output buffers, for the
 * integer data type. `ab_qc_quantize_info` carries what the reference's QcQuantizeInfo (QcQuantizeInfo.h:47-73) does; the
 * op modes have TensorQuantizerOpMode's values (DlQ/include/DlQuantization/TensorQuantizerOpFacade.h:48-54).
 * ---------------------------------------------------------------------------------------------------------- */
typedef enum
{
    AB_OP_UPDATE_STATS = 0,
    AB_OP_ONE_SHOT_QDQ = 1,
    AB_OP_QDQ          = 2,
    AB_OP_PASS_THROUGH = 3
} ab_op_mode;

typedef struct
{
    ab_stats_state* states; /* DEVICE: one record per encoding (replaces tensorQuantizerRef); may be NULL for QDQ / pass-through */
    ab_encoding* encodings; /* HOST: num_encodings entries; min / max / delta / offset are rewritten by the one-shot mode */
    int num_encodings;      /* 1, the channel count, or channels x blocks */
    int op_mode;            /* ab_op_mode; the one-shot mode switches itself to AB_OP_QDQ after running once */
    int quant_mode;         /* ab_quant_mode of the statistics */
    int use_symmetric_encoding;
    int enabled;
    int is_int_data_type;
    int use_per_channel_mode;
    int channel_axis;
    int block_axis;
    int block_size;         /* 0: plain per-channel */
} ab_qc_quantize_info;

/* bytes of DEVICE scratch ab_qc_quantize_op_compute needs for `num_encodings` encodings (the reference cudaMallocs per call) */
size_t ab_qc_quantize_op_workspace_bytes(int num_encodings);

/* shape / ndim: the input tensor's dimensions (contiguous). Statistics are supported per tensor, per channel when no
 * dimension precedes the channel axis, and per contiguous block; quantize-dequantize for every layout the reference handles.
 * Unlike the reference's CUDA op no stream synchronisation happens in updateStats mode. */
int ab_qc_quantize_op_compute(ab_qc_quantize_info* info, const void* in, void* out, const int64_t* shape, int ndim, int dtype,
                              void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * The entropy scheme: QuantizationMode::QUANTIZATION_ENTROPY = 5 of the reference's libpymo
 * (ModelOptimizations/PyModelOptimizations/PyModelOptimizations.cpp:148-156), EntropyEncodingAnalyzer<float>
 * (DlQ/src/EntropyEncodingAnalyzer.cpp). `state` is an ab_stats_state record initialised by ab_stats_reset and used with
 * these three functions only.
 *  ab_entropy_update           : updateStats (:80-96) -> updateTensorHistogram (DlQ/src/math_functions.cpp:440-560): min / max
 *                                of the tensor, range growth with redistribution of the older counts, binning. Three launches
 *                                on `stream`, no host synchronisation (the reference's GPU build copies the tensor to the
 *                                host and bins it there).
 *  ab_entropy_compute_encoding : computeEncoding (:98-143) -> _optimizeKL (:221-428) -> getComputedEncodings. Reads the
 *                                512-bin histogram back (synchronises `stream`) and runs the KL search on the HOST: the
 *                                result is defined by the C library's log (csrc/entropy_math.h). `out` is HOST memory.
 *  ab_entropy_histogram        : the raw TensorProfilingParams {histogram[512], min, max}; info[3] = {initialized,
 *                                stats_updated, iterations} (may be NULL). The reference's own getStatsHistogram for this
 *                                scheme (:54-78) trips its size assertion, so there is no PDF view to mirror.
 * ------------------------------------------------------------------------------------------------------------ */
#define AB_QUANTIZATION_ENTROPY 5
int ab_entropy_update(const void* in, int64_t count, int dtype, ab_stats_state* state, void* stream);
int ab_entropy_compute_encoding(const ab_stats_state* state, int bw, int use_symmetric, int use_strict_symmetric,
                                int use_unsigned_symmetric, ab_encoding* out, void* stream);
int ab_entropy_histogram(const ab_stats_state* state, double* hist512, double* min_max, int* info, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AIMET_B200_H_ */
