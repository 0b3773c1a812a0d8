/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's quantization-simulation
 * hot path, in plain C. Nothing under aimet_b200/ may include, link or call this.
 *
 * Parity status: PINNED. tests/test_oracle_pin.py checks every function here against
 *   (a) the reference's own known-answer vectors (the .cpp files under DlQuantization/test, cited per test), and
 *   (b) the reference's unmodified C++ compiled from /root/reference into oracle/_ref/libaimet_ref.so
 *       (oracle/Makefile, oracle/ref_shim.cpp) on seeded random inputs, bit for bit, and
 *   (c) the committed golden fixtures under tests/golden/ generated from (b).
 *
 * All file:line citations are into /root/reference/ModelOptimizations/DlQuantization unless noted.
 */
#ifndef QSIM_ORACLE_H_
#define QSIM_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QO_PDF_SIZE 512 /* src/math_functions.hpp:80 */

/* include/DlQuantization/Quantization.hpp:113-120 */
typedef struct
{
    double min, max, delta, offset;
    int bw;
} qo_encoding;

/* src/math_functions.hpp:59-67 (PDF) + the analyzer's _statsUpdated flag (TfEnhancedEncodingAnalyzer.h) */
typedef struct
{
    int initialized;   /* xLeft.size() != 0 */
    int stats_updated; /* _statsUpdated */
    int iterations;
    double x_left[QO_PDF_SIZE];
    double pdf[QO_PDF_SIZE];
} qo_tfe_state;

/* src/TfEncodingAnalyzer.h:85-93 */
typedef struct
{
    int stats_updated;
    double min, max;
} qo_tf_state;

/* ---- encodings from min/max (host double math) ---- */
void qo_gate_min_max(double* enc_min, double* enc_max);                           /* src/quantization_utils.cpp:145-156 */
void qo_fill_encoding_info(int bw, double enc_min, double enc_max, qo_encoding* e); /* src/TensorQuantizationSim.cpp:63-92 */
qo_encoding qo_tf_encoding(int bw, double mn, double mx, int sym, int strict, int unsigned_sym); /* src/quantization_utils.cpp:58-143 */
/* src/TensorQuantizer.cpp:327-343 ; returns 0 ok, 1 = reference throws */
int qo_partial_encoding(int bw, qo_encoding* e, int sym, int unsigned_sym, int strict);

/* ---- element-wise kernels ---- */
void qo_qdq(const float* in, size_t n, float* out, const qo_encoding* e);          /* src/trim_functions.cpp:140-182 */
void qo_qdq_tensor(const float* in, size_t n, float* out, double enc_min, double enc_max, int bw); /* TensorQuantizationSim.cpp:104-114 */
int64_t qo_quantize_packed(const float* in, size_t n, uint8_t* out, double enc_min, double enc_max, int bw,
                           int shift_to_signed);
void qo_quantize_tensor(const float* in, size_t n, float* out, double enc_min, double enc_max, int bw,
                        int shift_to_signed);                                       /* TensorQuantizationSim.cpp:116-126; trim_functions.cpp:202-218 */
/* TrainingExtensions/torch/src/AimetTensorQuantizer.cpp:236-299 (torch fp32 CPU ops restated) */
void qo_per_channel_prepare(const double* enc_min, const double* enc_max, int num_channel, int bw, float* o_min,
                            float* o_max, float* o_delta, float* o_offset);
void qo_qdq_per_channel(const float* in, size_t num_channel, size_t num_element, size_t num_element_per_channel,
                        float* out, const float* enc_min, const float* enc_max, const float* enc_delta,
                        const float* enc_offset);                                   /* src/trim_functions.cpp:697-709 */
/* src/trim_functions.cpp:633-662 (quantizeDequantizeBroadcastCpu) */
void qo_qdq_broadcast(const float* in, float* out, int64_t num_element, int64_t num_dims, const int64_t* input_strides,
                      const int64_t* encoding_strides, const float* enc_min, const float* enc_max, const float* enc_delta,
                      const float* enc_offset);
/* TrainingExtensions/torch/src/python/aimet_torch/v1/quantsim_straight_through_grad.py:91-118 */
void qo_ste_bwd(const float* x, const float* grad, size_t n, float enc_min, float enc_max, float* grad_in);
void qo_ste_bwd_per_channel(const float* x, const float* grad, size_t num_channel, size_t num_element,
                            size_t num_element_per_channel, const float* enc_min, const float* enc_max,
                            float* grad_in);
/* bf16 round trips used by the Python hosts (tensor.to(float32) ... .to(bfloat16)); RNE */
float qo_bf16_to_f32(uint16_t v);
uint16_t qo_f32_to_bf16(float v);

/* ---- statistics ---- */
float qo_get_min(const float* data, size_t n);                                      /* src/math_functions.cpp:338-347 */
float qo_get_max(const float* data, size_t n);                                      /* src/math_functions.cpp:327-336 */
void qo_histogram(const float* data, size_t n, uint32_t* hist, float bucket_size, float pdf_offset); /* math_functions.cpp:367-384 */
/* the (bucket_size, pdf_offset) pair UpdatePdf derives from an initialised PDF: math_functions.cpp:264-268 */
void qo_tfe_bucket_params(const qo_tfe_state* s, float* bucket_size, float* pdf_offset);

void qo_tf_init(qo_tf_state* s);
void qo_tf_update(qo_tf_state* s, const float* data, size_t n);                     /* src/TfEncodingAnalyzer.cpp:60-71 */
qo_encoding qo_tf_compute(const qo_tf_state* s, int bw, int sym, int strict, int unsigned_sym); /* :81-101 */

void qo_tfe_init(qo_tfe_state* s);
void qo_tfe_init_pdf(qo_tfe_state* s, float min_val, float max_val);                /* src/math_functions.cpp:207-241 */
void qo_tfe_update(qo_tfe_state* s, const float* data, size_t n);                   /* src/math_functions.cpp:243-288 */
/* fold one batch's integer histogram (already binned with this state's range) into the running PDF: :279-287 */
void qo_tfe_fold_histogram(qo_tfe_state* s, const uint32_t* hist, size_t cnt);
qo_encoding qo_tfe_compute(const qo_tfe_state* s, int bw, int sym, int strict, int unsigned_sym); /* src/TfEnhancedEncodingAnalyzer.cpp:79-113,358-397 */
/* percentile calibration on the same statistics: src/PercentileEncodingAnalyzer.cpp:77-196 */
qo_encoding qo_percentile_compute(const qo_tfe_state* s, float percentile, int bw, int sym, int strict, int unsigned_sym);
/* MSE calibration on the same statistics: src/MseEncodingAnalyzer.cpp:77-285 */
qo_encoding qo_mse_compute(const qo_tfe_state* s, int bw, int sym, int strict, int unsigned_sym);
/* cost of one candidate: TfEnhancedEncodingAnalyzer.cpp:294-355 */
double qo_tfe_cost(const qo_tfe_state* s, int bw, float delta, int offset);
/* candidate list (delta[], offset[]) in the reference's order; returns count (<= 358): :178-253 */
int qo_tfe_candidates(const qo_tfe_state* s, int bw, int sym, int strict, int unsigned_sym, float* deltas,
                      int* offsets, float* num_steps_out);

/* ---- entropy scheme: src/EntropyEncodingAnalyzer.cpp, TensorProfilingParams src/math_functions.hpp:71-77 ---- */
typedef struct
{
    int initialized;   /* histogram.size() != 0 */
    int stats_updated; /* _statsUpdated */
    int iterations;
    double min, max;
    double histogram[QO_PDF_SIZE];
} qo_entropy_state;
void qo_entropy_init(qo_entropy_state* s);
void qo_entropy_update(qo_entropy_state* s, const float* data, size_t n);   /* :80-96 -> math_functions.cpp:472-560 */
qo_encoding qo_entropy_compute(const qo_entropy_state* s, int bw, int sym, int strict, int unsigned_sym); /* :98-143 */
/* rescaleHistogram, math_functions.cpp:562-640 (512 bins) */
void qo_rescale_histogram(const double* src, double src_min, double src_max, double dst_min, double dst_max, double* dst);

#ifdef __cplusplus
}
#endif
#endif /* QSIM_ORACLE_H_ */
