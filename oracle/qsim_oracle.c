/* TEST INFRASTRUCTURE ONLY -- see qsim_oracle.h. Plain-C restatement of the reference's CPU path.
 * Build: gcc -std=c11 -O3 -ffp-contract=off (the reference is built in ISO C++ mode, which disables
 * FMA contraction, and without -march, so every expression below is evaluated operation by operation
 * in the type C's promotion rules give it -- that typing is the whole point of this file).
 */
#include "qsim_oracle.h"

#include <float.h>
#include <limits.h>
#include <math.h>
#include <string.h>

#define QO_EPSILON 1e-5    /* src/quantization_utils.hpp:51 */
#define QO_MIN_RANGE 0.01  /* src/TfEncodingAnalyzer.h:79, src/TfEnhancedEncodingAnalyzer.h:105 */
#define QO_GAMMA 3.0f      /* src/TfEnhancedEncodingAnalyzer.h:102 (DTYPE = float) */

/* std::min / std::max as libstdc++ defines them: min(a,b) = (b<a)?b:a ; max(a,b) = (a<b)?b:a */
static inline double dmin(double a, double b) { return (b < a) ? b : a; }
static inline double dmax(double a, double b) { return (a < b) ? b : a; }
static inline float fminx(float a, float b) { return (b < a) ? b : a; }
static inline float fmaxx(float a, float b) { return (a < b) ? b : a; }
static inline int imin(int a, int b) { return (b < a) ? b : a; }
static inline int imax(int a, int b) { return (a < b) ? b : a; }

/* float/double -> int the way x86-64 (cvttss2si / cvttsd2si) does it: NaN and out-of-range give INT_MIN.
 * The reference relies on this (it is UB in C++), e.g. histogram samples far outside the range. */
static inline int f2i(float v)
{
    if (!(v > -2147483904.0f && v < 2147483648.0f))
        return INT_MIN;
    return (int) v;
}
static inline int d2i(double v)
{
    if (!(v > -2147483649.0 && v < 2147483648.0))
        return INT_MIN;
    return (int) v;
}
/* int + int with wrap-around (what the compiled reference does on overflow) */
static inline int iadd_wrap(int a, int b) { return (int) ((unsigned) a + (unsigned) b); }

/* ------------------------------------------------------------------------------------------------ */
/* encodings from min / max                                                                           */
/* ------------------------------------------------------------------------------------------------ */

void qo_gate_min_max(double* enc_min, double* enc_max)
{
    *enc_min = dmin(*enc_min, 0.0);
    *enc_max = dmax(*enc_max, 0.0);
    *enc_max = dmax(*enc_max, *enc_min + QO_EPSILON);
}

void qo_fill_encoding_info(int bw, double enc_min, double enc_max, qo_encoding* e)
{
    e->bw  = (uint8_t) bw;
    e->min = enc_min;
    e->max = enc_max;
    qo_gate_min_max(&e->min, &e->max);
    double num_steps = pow(2, (uint8_t) bw) - 1;
    if (e->min == -e->max)
        num_steps -= 1;
    e->delta  = (e->max - e->min) / num_steps;   /* trim_functions.cpp:61-65 */
    e->offset = round(e->min / e->delta);        /* trim_functions.cpp:68-73 */
    e->min    = e->offset * e->delta;
    e->max    = e->delta * num_steps + e->min;
}

qo_encoding qo_tf_encoding(int bw_in, double mn, double mx, int sym, int strict, int unsigned_sym)
{
    qo_encoding e;
    uint8_t bw       = (uint8_t) bw_in;
    double num_steps = pow(2, bw) - 1;
    if (sym && strict)
        num_steps -= 1;
    e.bw = bw;
    if (isinf(mn))
        mn = -(double) FLT_MAX;
    if (isinf(mx))
        mx = (double) FLT_MAX;

    if (sym && ((mn < 0.0) || !unsigned_sym))
    {
        mx                         = dmax(fabs(mx), fabs(mn));
        unsigned int num_pos_steps = (unsigned int) floor(num_steps / 2);
        e.delta                    = mx / num_pos_steps;
        e.offset                   = -ceil(num_steps / 2);
        e.min                      = dmax(e.offset * e.delta, -(double) FLT_MAX);
        e.max                      = dmin(e.delta * num_pos_steps, (double) FLT_MAX);
    }
    else
    {
        e.delta = (mx - mn) / num_steps;
        if (mn < 0 && mx > 0)
        {
            double b_zero = round(-mn / e.delta);
            b_zero        = dmin(num_steps, dmax(0.0, b_zero));
            e.offset      = -b_zero;
        }
        else
        {
            e.offset = round(mn / e.delta);
            e.min    = mn;
            e.max    = mx;
            return e;
        }
        if (e.delta * e.offset >= -(double) FLT_MAX && e.delta * e.offset <= (double) FLT_MAX)
            e.min = e.delta * e.offset;
        else
            e.min = -(double) FLT_MAX;
        e.max = mx - mn + e.min;
        if (e.max > (double) FLT_MAX)
            e.max = (double) FLT_MAX;
    }
    return e;
}

/* src/quantization_utils.cpp:158-205 */
static int min_max_from_delta_offset(int bw_in, qo_encoding* e, int sym, int unsigned_sym, int strict)
{
    uint8_t bw = (uint8_t) bw_in;
    if (e->bw == 0)
        return 1;
    if (e->min != 0 && e->max != 0)
        return 1;
    if (e->delta == 0 && e->offset > 0)
        return 1;
    double num_steps = pow(2, bw) - 1;
    if (sym && strict)
        num_steps -= 1;
    e->min = e->offset * e->delta;
    if (sym && ((e->min < 0.0) || !unsigned_sym))
    {
        double num_pos_steps = floor(num_steps / 2);
        e->max               = e->delta * num_pos_steps;
    }
    else
    {
        e->max = e->delta * num_steps + e->min;
    }
    if (e->max - e->min < QO_EPSILON)
        qo_gate_min_max(&e->min, &e->max);
    return 0;
}

/* src/quantization_utils.cpp:207-228 */
static int delta_offset_from_min_max(int bw_in, qo_encoding* e, int sym, int unsigned_sym, int strict)
{
    qo_encoding orig = *e;
    if (e->bw == 0)
        return 1;
    if (orig.delta != 0 && orig.offset != 0)
        return 1;
    *e     = qo_tf_encoding(bw_in, e->min, e->max, sym, strict, unsigned_sym);
    e->min = orig.min;
    e->max = orig.max;
    return 0;
}

int qo_partial_encoding(int bw, qo_encoding* e, int sym, int unsigned_sym, int strict)
{
    if (e->min == 0 && e->max == 0)
        return min_max_from_delta_offset(bw, e, sym, unsigned_sym, strict);
    else if (e->delta == 0)
        return delta_offset_from_min_max(bw, e, sym, unsigned_sym, strict);
    return 1;
}

/* ------------------------------------------------------------------------------------------------ */
/* element-wise kernels (DTYPE = float: every operation below is a float operation)                   */
/* ------------------------------------------------------------------------------------------------ */

/* src/trim_functions.cpp:140-166, ROUND_NEAREST only (stochastic rounding is rand()-seeded: not reproducible) */
static inline float quantize_value(float x, float e_min, float e_max, float e_delta, float e_offset)
{
    float v = fmaxf(fminf(x, e_max), e_min);
    v       = v / e_delta - e_offset;
    return roundf(v);
}

void qo_qdq(const float* in, size_t n, float* out, const qo_encoding* e)
{
    /* doubles narrow to float at the call (trim_functions.cpp:178) */
    const float e_min = (float) e->min, e_max = (float) e->max, e_delta = (float) e->delta,
                e_offset = (float) e->offset;
    for (size_t i = 0; i < n; ++i)
    {
        float q = quantize_value(in[i], e_min, e_max, e_delta, e_offset);
        out[i]  = e_delta * (q + e_offset);   /* trim_functions.cpp:168-172 */
    }
}

void qo_qdq_tensor(const float* in, size_t n, float* out, double enc_min, double enc_max, int bw)
{
    qo_encoding e;
    qo_fill_encoding_info(bw, enc_min, enc_max, &e);
    qo_qdq(in, n, out, &e);
}

void qo_quantize_tensor(const float* in, size_t n, float* out, double enc_min, double enc_max, int bw,
                        int shift_to_signed)
{
    qo_encoding e;
    qo_fill_encoding_info(bw, enc_min, enc_max, &e);
    const float e_min = (float) e.min, e_max = (float) e.max, e_delta = (float) e.delta, e_offset = (float) e.offset;
    unsigned int shift = 0;
    if (shift_to_signed)
        shift = (unsigned int) pow(2, e.bw - 1);
    for (size_t i = 0; i < n; ++i)
    {
        float q = quantize_value(in[i], e_min, e_max, e_delta, e_offset);
        out[i]  = q - (float) shift;   /* out[i] -= shift : unsigned -> float, float subtract */
    }
}

/* quantizeTensorPacked + quantizeToFxpPackedCpu (TensorQuantizationSim.cpp:128-139, trim_functions.cpp:221-388), nearest
 * rounding. All arithmetic in DOUBLE on the double encoding (unlike the float kernels above). One value per byte below 8 bit
 * (the reference's sub-byte packing is compiled out), uint16 / uint32 above; shiftToSigned subtracts 2^(bw-1) - 1 (NOT
 * 2^(bw-1) as quantizeToFxp does) and stores int8 / int16 / int32. Returns bytes written, -1 for an unsupported bitwidth.
 * The float -> integer casts are x86's cvttsd2si on values already clamped into range, except NaN inputs, which the
 * reference lets through its std::min / std::max chain (they return the first argument when a comparison with NaN is
 * false) and then casts -- reproduced as the instruction behaves: the "integer indefinite" value, truncated. */
static int64_t cvtt_i64(double v)
{
    if (!(v == v) || v >= 9223372036854775808.0 || v < -9223372036854775808.0)
        return INT64_MIN;
    return (int64_t) v;
}
static int32_t cvtt_i32(double v)
{
    if (!(v == v) || v >= 2147483648.0 || v < -2147483649.0)
        return INT32_MIN;
    return (int32_t) v;
}
static double std_min(double a, double b) { return (b < a) ? b : a; }
static double std_max(double a, double b) { return (a < b) ? b : a; }

int64_t qo_quantize_packed(const float* in, size_t n, uint8_t* out, double enc_min, double enc_max, int bw,
                           int shift_to_signed)
{
    qo_encoding e;
    qo_fill_encoding_info(bw, enc_min, enc_max, &e);
    if (!(bw == 1 || bw == 2 || bw == 4 || bw == 8 || bw == 16 || bw == 32))
        return -1;
    const int bytes_per = (bw > 8 ? bw : 8) / 8;
    for (size_t i = 0; i < n; ++i)
    {
        double q = std_max(std_min((double) in[i], e.max), e.min);
        q        = q / e.delta - e.offset;
        q        = round(q);
        if (!shift_to_signed)
        {
            if (bw < 8)
            {
                const uint8_t shr = (uint8_t) cvtt_i32(q);                      /* (uint8_t) data_quantized */
                out[i]            = (uint8_t) cvtt_i32(std_max(std_min((double) shr, pow(2, bw) - 1), 0.0));
            }
            else if (bw == 8)
                out[i] = (uint8_t) cvtt_i32(std_max(std_min(q, 255.0), 0.0));
            else if (bw == 16)
                ((uint16_t*) out)[i] = (uint16_t) cvtt_i32(std_max(std_min(q, 65535.0), 0.0));
            else
                ((uint32_t*) out)[i] = (uint32_t) cvtt_i64(std_max(std_min(q, 4294967295.0), 0.0));
        }
        else
        {
            q -= pow(2, bw - 1) - 1;
            if (bw < 8)
                ((int8_t*) out)[i] = (int8_t) ((int8_t) cvtt_i32(q) & (int8_t) (pow(2, bw) - 1));
            else if (bw == 8)
                ((int8_t*) out)[i] = (int8_t) cvtt_i32(std_max(std_min(q, 127.0), -128.0));
            else if (bw == 16)
                ((int16_t*) out)[i] = (int16_t) cvtt_i32(std_max(std_min(q, 32767.0), -32768.0));
            else
                ((int32_t*) out)[i] = cvtt_i32(std_max(std_min(q, 2147483647.0), -2147483648.0));
        }
    }
    return (int64_t) n * bytes_per;
}

void qo_per_channel_prepare(const double* enc_min, const double* enc_max, int num_channel, int bw, float* o_min,
                            float* o_max, float* o_delta, float* o_offset)
{
    /* AimetTensorQuantizer.cpp:286-294: step count decided from channel 0 only, in double */
    double num_steps = pow(2, bw) - 1;
    if (enc_min[0] == -enc_max[0])
        num_steps -= 1;
    const float steps_f = (float) num_steps;   /* tensor / Scalar: the scalar is cast to the tensor's dtype */
    const float eps_f   = (float) 1e-5;
    for (int c = 0; c < num_channel; ++c)
    {
        float mn = (float) enc_min[c];   /* std::vector<float> filled from doubles: :272-277 */
        float mx = (float) enc_max[c];
        /* gateMinMaxTensor :236-242 : torch.minimum / torch.maximum on fp32 */
        mn = (mn < 0.0f) ? mn : 0.0f;
        mx = (mx > 0.0f) ? mx : 0.0f;
        float lo = mn + eps_f;
        mx       = (mx > lo) ? mx : lo;
        float d  = (mx - mn) / steps_f;          /* :244-248 */
        float o  = nearbyintf(mn / d);           /* at::round = half-to-even, :250-254 */
        o_min[c] = mn, o_max[c] = mx, o_delta[c] = d, o_offset[c] = o;
    }
}

void qo_qdq_per_channel(const float* in, size_t num_channel, size_t num_element, size_t num_element_per_channel,
                        float* out, const float* enc_min, const float* enc_max, const float* enc_delta,
                        const float* enc_offset)
{
    for (size_t i = 0; i < num_element; ++i)
    {
        size_t c = (i / num_element_per_channel) % num_channel;
        float q  = quantize_value(in[i], enc_min[c], enc_max[c], enc_delta[c], enc_offset[c]);
        out[i]   = enc_delta[c] * (q + enc_offset[c]);
    }
}

void qo_ste_bwd(const float* x, const float* grad, size_t n, float enc_min, float enc_max, float* grad_in)
{
    /* mask = (min <= x) & (x <= max); grad * mask  (bool promotes to 1.0f / 0.0f, so inf * 0 = NaN survives) */
    for (size_t i = 0; i < n; ++i)
    {
        float m    = (enc_min <= x[i] && x[i] <= enc_max) ? 1.0f : 0.0f;
        grad_in[i] = grad[i] * m;
    }
}

void qo_ste_bwd_per_channel(const float* x, const float* grad, size_t num_channel, size_t num_element,
                            size_t num_element_per_channel, const float* enc_min, const float* enc_max,
                            float* grad_in)
{
    for (size_t i = 0; i < num_element; ++i)
    {
        size_t c   = (i / num_element_per_channel) % num_channel;
        float m    = (enc_min[c] <= x[i] && x[i] <= enc_max[c]) ? 1.0f : 0.0f;
        grad_in[i] = grad[i] * m;
    }
}

float qo_bf16_to_f32(uint16_t v)
{
    uint32_t u = ((uint32_t) v) << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

uint16_t qo_f32_to_bf16(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u)   /* NaN: torch returns the canonical quiet NaN 0x7fc0 */
        return 0x7fc0;
    uint32_t lsb = (u >> 16) & 1u;
    u += 0x7fffu + lsb;
    return (uint16_t) (u >> 16);
}

/* ------------------------------------------------------------------------------------------------ */
/* statistics                                                                                         */
/* ------------------------------------------------------------------------------------------------ */

float qo_get_max(const float* data, size_t n)
{
    float val = (float) -DBL_MAX;   /* = -inf */
    for (size_t i = 0; i < n; ++i)
        val = fmaxx(val, data[i]);
    return val;
}

float qo_get_min(const float* data, size_t n)
{
    float val = (float) DBL_MAX;   /* = +inf */
    for (size_t i = 0; i < n; ++i)
        val = fminx(val, data[i]);
    return val;
}

void qo_histogram(const float* data, size_t n, uint32_t* hist, float bucket_size, float pdf_offset)
{
    for (size_t i = 0; i < n; ++i)
    {
        int index = f2i(roundf(data[i] / bucket_size - pdf_offset));
        if (index >= 0 && index < QO_PDF_SIZE)
            hist[index] += 1;
    }
}

void qo_tf_init(qo_tf_state* s)
{
    s->stats_updated = 0;
    s->min           = DBL_MAX;
    s->max           = -DBL_MAX;
}

void qo_tf_update(qo_tf_state* s, const float* data, size_t n)
{
    s->stats_updated = 1;
    double cur_min   = (double) qo_get_min(data, n);
    double cur_max   = (double) qo_get_max(data, n);
    s->min           = dmin(s->min, cur_min);
    s->max           = dmax(s->max, cur_max);
}

qo_encoding qo_tf_compute(const qo_tf_state* s, int bw, int sym, int strict, int unsigned_sym)
{
    double new_min = dmin(0.0, s->min);
    double new_max = dmax(0.0, s->max);
    new_max        = dmax(new_max, new_min + QO_MIN_RANGE);
    return qo_tf_encoding(bw, new_min, new_max, sym, strict, unsigned_sym);
}

void qo_tfe_init(qo_tfe_state* s)
{
    memset(s, 0, sizeof(*s));
}

void qo_tfe_init_pdf(qo_tfe_state* s, float min_val, float max_val)
{
    if (min_val == max_val)
        max_val = fmaxx(max_val, min_val + (float) 0.01);
    float center = (max_val + min_val) / 2;
    min_val      = fmaxx(-FLT_MAX, center - 3 * (center - min_val));
    max_val      = fminx(FLT_MAX, center + 3 * (max_val - center));
    double bucket_size = ((double) max_val - (double) min_val) / QO_PDF_SIZE;   /* signed_vals == true */
    for (int i = 0; i < QO_PDF_SIZE; ++i)
    {
        s->x_left[i] = min_val + i * bucket_size;
        s->pdf[i]    = 0.0;
    }
    s->iterations  = 0;
    s->initialized = 1;
}

void qo_tfe_bucket_params(const qo_tfe_state* s, float* bucket_size, float* pdf_offset)
{
    float b      = (float) (s->x_left[1] - s->x_left[0]);
    float mn     = (float) s->x_left[0];
    *bucket_size = b;
    *pdf_offset  = mn / b;
}

void qo_tfe_fold_histogram(qo_tfe_state* s, const uint32_t* hist, size_t cnt_in)
{
    int cnt = (int) cnt_in;   /* UpdatePdf takes `int cnt` */
    for (int i = 0; i < QO_PDF_SIZE; ++i)
    {
        double prob = (double) hist[i] / (double) cnt;
        s->pdf[i]   = (s->pdf[i] * s->iterations + prob) / (s->iterations + 1);
    }
    s->iterations++;
}

void qo_tfe_update(qo_tfe_state* s, const float* data, size_t n)
{
    s->stats_updated = 1;
    if (!s->initialized)
    {
        float mn = qo_get_min(data, n);
        float mx = qo_get_max(data, n);
        if (mn == 0 && mx == 0)
            return;
        qo_tfe_init_pdf(s, mn, mx);
    }
    float bucket_size, pdf_offset;
    qo_tfe_bucket_params(s, &bucket_size, &pdf_offset);
    uint32_t hist[QO_PDF_SIZE];
    memset(hist, 0, sizeof(hist));
    qo_histogram(data, n, hist, bucket_size, pdf_offset);
    qo_tfe_fold_histogram(s, hist, n);
}

double qo_tfe_cost(const qo_tfe_state* s, int bw, float delta, int offset)
{
    float min_val   = delta * offset;
    float step_size = (float) (pow(2, bw) - 1);
    float max_val   = delta * (offset + step_size);
    float pdf_start = (float) s->x_left[0];
    double pdf_step = s->x_left[1] - s->x_left[0];
    int min_ind     = d2i(floor((min_val - pdf_start) / pdf_step));
    min_ind         = imin(imax(0, min_ind), QO_PDF_SIZE - 1);
    int max_ind     = d2i(floor((max_val - pdf_start) / pdf_step));
    max_ind         = imin(imax(0, max_ind), QO_PDF_SIZE - 1);

    double sat_bottom  = 0;
    float min_mid      = (float) (pdf_start + (min_ind * pdf_step) + pdf_step / 2);
    for (int i = 0; i < min_ind; ++i)
    {
        double mid = pdf_start + i * pdf_step + pdf_step / 2;
        double d   = mid - min_mid;
        sat_bottom += s->pdf[i] * (d * d);
    }
    double sat_top = 0;
    float max_mid  = (float) (pdf_start + (max_ind * pdf_step) + pdf_step / 2);
    for (int i = max_ind; i < QO_PDF_SIZE; ++i)
    {
        double mid = pdf_start + i * pdf_step + pdf_step / 2;
        double d   = mid - max_mid;
        sat_top += s->pdf[i] * (d * d);
    }
    double quant_cost = 0;
    for (int i = min_ind; i < max_ind; ++i)
    {
        float float_val   = (float) (pdf_start + i * pdf_step + pdf_step / 2);
        int quantized     = f2i(roundf(float_val / delta - offset));
        float dequantized = delta * iadd_wrap(quantized, offset);
        double d          = (double) (float_val - dequantized);
        quant_cost += s->pdf[i] * (d * d);
    }
    double sqnr = QO_GAMMA * (sat_bottom + sat_top) + quant_cost;
    return dmin(sqnr, DBL_MAX);
}

/* TfEnhancedEncodingAnalyzer.cpp:256-291 */
static void tfe_range(const qo_tfe_state* s, float* o_min, float* o_max)
{
    float min_val = (float) s->x_left[0];
    float max_val = (float) s->x_left[QO_PDF_SIZE - 1];
    for (int i = 0; i < QO_PDF_SIZE; ++i)
        if (s->pdf[i] > 0)
        {
            min_val = (float) s->x_left[i];
            break;
        }
    for (int i = QO_PDF_SIZE - 1; i > 0; --i)
        if (s->pdf[i] > 0)
        {
            max_val = (float) s->x_left[i];
            break;
        }
    min_val = fminx(min_val, 0.0f);
    max_val = fmaxx(max_val, 0.0f);
    max_val = fmaxx(max_val, min_val + (float) QO_MIN_RANGE);
    *o_min = min_val, *o_max = max_val;
}

/* TfEnhancedEncodingAnalyzer.cpp:146-175 */
static int clamp_to_observed(float obs_min, float obs_max, float num_steps, float* test_delta, int* test_offset)
{
    float t_min = fmaxx(*test_delta * *test_offset, -FLT_MAX);
    float t_max = fminx(*test_delta * (*test_offset + num_steps), FLT_MAX);
    if ((t_min < obs_min) && (t_max > obs_max))
        return 0;
    t_min = fmaxx(obs_min, t_min);
    t_max = fminx(obs_max, t_max);
    if (t_min == t_max)
        return 0;
    *test_delta  = (float) (((double) t_max - t_min) / num_steps);
    *test_offset = f2i(roundf(t_min / *test_delta));
    return 1;
}

int qo_tfe_candidates(const qo_tfe_state* s, int bw, int sym, int strict, int unsigned_sym, float* deltas,
                      int* offsets, float* num_steps_out)
{
    float min_val, max_val;
    tfe_range(s, &min_val, &max_val);
    float num_steps = (float) (pow(2, bw) - 1);
    int n           = 0;
    if (sym)
    {
        if (strict)
            num_steps -= 1;
        /* _pickTestCandidatesSymmetric :217-253 */
        float delta_max;
        int test_offset;
        if ((min_val == 0.0) && unsigned_sym)
        {
            delta_max   = max_val / num_steps;
            test_offset = 0;
        }
        else
        {
            float abs_max = fmaxx(fabsf(max_val), fabsf(min_val));
            delta_max     = (float) (abs_max / (num_steps / 2.0));
            test_offset   = f2i(floorf(-num_steps / 2));
        }
        for (float f = (float) (1.0 / 100); f <= 1 + 1.0 / 100; f = (float) (f + 1.0 / 100))
        {
            deltas[n]  = f * delta_max;
            offsets[n] = test_offset;
            ++n;
        }
    }
    else
    {
        /* _pickTestCandidatesAsymmetric :178-214 */
        float obs_min    = min_val, obs_max = max_val;
        float obs_delta  = (float) (((double) obs_max - (double) obs_min) / num_steps);
        int obs_offset   = f2i(roundf(obs_min / obs_delta));
        obs_min          = fmaxx(obs_delta * obs_offset, -FLT_MAX);
        obs_max          = fminx(obs_delta * (obs_offset + num_steps), FLT_MAX);
        float delta_max  = obs_delta;
        for (float f = (float) (1.0 / 16); f <= 1 + 1.0 / 16; f = (float) (f + 1.0 / 16))
        {
            for (int i = 0; i <= 20; ++i)
            {
                float test_delta = f * delta_max;
                int test_offset  = d2i(-num_steps + num_steps / 20.0 * i);
                if (!clamp_to_observed(obs_min, obs_max, num_steps, &test_delta, &test_offset))
                    continue;
                deltas[n]  = test_delta;
                offsets[n] = test_offset;
                ++n;
            }
        }
        deltas[n]  = obs_delta;
        offsets[n] = obs_offset;
        ++n;
    }
    *num_steps_out = num_steps;
    return n;
}

qo_encoding qo_tfe_compute(const qo_tfe_state* s, int bw_in, int sym, int strict, int unsigned_sym)
{
    qo_encoding e = {0, 0, 0, 0, 0};
    uint8_t bw    = (uint8_t) bw_in;
    if (!s->initialized)
    {
        if (s->stats_updated)
        {
            /* all-zero data so far: TfEnhancedEncodingAnalyzer.cpp:85-100 */
            float num_steps = (float) (pow(2, bw) - 1);
            e.min           = -1;
            e.max           = 1;
            e.delta         = (e.max - e.min) / (int) num_steps;
            e.offset        = floor(e.min / e.delta);
            e.min           = e.offset * e.delta;
            e.max           = e.min + (int) num_steps * e.delta;
            e.bw            = bw;
        }
        return e;
    }
    float deltas[360];
    int offsets[360];
    float num_steps;
    int n = qo_tfe_candidates(s, bw, sym, strict, unsigned_sym, deltas, offsets, &num_steps);

    /* _findBestCandidate :115-144 */
    float best_delta = -1;
    int best_offset  = -1;
    double best_cost = DBL_MAX;
    for (int k = 0; k < n; ++k)
    {
        double cost = qo_tfe_cost(s, bw, deltas[k], offsets[k]);
        if (cost < best_cost)
        {
            best_cost   = cost;
            best_delta  = deltas[k];
            best_offset = offsets[k];
        }
    }
    float best_min = fmaxx(best_delta * best_offset, -FLT_MAX);
    float best_max = fminx(best_delta * (best_offset + num_steps), FLT_MAX);
    e.delta        = best_delta;
    e.offset       = best_offset;
    e.bw           = bw;
    e.min          = best_min;
    e.max          = best_max;
    return e;
}

/* ---- percentile calibration: src/PercentileEncodingAnalyzer.cpp:77-196 (DTYPE = float). The statistics are the
 * tf_enhanced ones (UpdatePdf, :69-75), so a qo_tfe_state carries them. ---- */
static void percentile_range(const qo_tfe_state* s, float percentile, float* o_min, float* o_max)
{
    float min_val, max_val;
    tfe_range(s, &min_val, &max_val); /* findOriginalRange, math_functions.cpp:404-436: same function */
    if (percentile == 100.0f)
    {
        *o_min = min_val, *o_max = max_val;
        return;
    }
    const float bin_width = (float) (s->x_left[1] - s->x_left[0]);
    float hist_min        = (float) s->x_left[0];
    float hist_max        = (float) (s->x_left[QO_PDF_SIZE - 1] + bin_width);
    float p_min = hist_min, p_max = hist_max;
    double cdf[QO_PDF_SIZE];
    memcpy(cdf, s->pdf, sizeof(cdf));
    for (int i = 1; i < QO_PDF_SIZE; i++)
        cdf[i] += cdf[i - 1];
    float left = 1 - percentile / 100;
    for (int i = 0; i < QO_PDF_SIZE; i++)
        if (cdf[i] >= left)
        {
            p_min = (float) s->x_left[i];
            break;
        }
    float right = percentile / 100;
    for (int i = QO_PDF_SIZE - 1; i >= 0; i--)
        if (cdf[i] < right && s->x_left[i] < max_val)
        {
            p_max = (float) (s->x_left[i] + bin_width);
            break;
        }
    if (p_min == p_max)
        p_max += bin_width;
    *o_min = p_min, *o_max = p_max;
}

qo_encoding qo_percentile_compute(const qo_tfe_state* s, float percentile, int bw_in, int sym, int strict,
                                  int unsigned_sym)
{
    qo_encoding e   = {0, 0, 0, 0, 0};
    uint8_t bw      = (uint8_t) bw_in;
    float num_steps = (float) (pow(2, bw) - 1);
    if (sym && strict)
        num_steps -= 1;
    if (!s->initialized)
    {
        if (s->stats_updated)
        {
            e.min    = -1;
            e.max    = 1;
            e.delta  = (e.max - e.min) / (int) num_steps;
            e.offset = floor(e.min / e.delta);
            e.min    = e.offset * e.delta;
            e.max    = e.min + (int) num_steps * e.delta;
            e.bw     = bw;
        }
        return e;
    }
    float a_min, a_max;
    percentile_range(s, percentile, &a_min, &a_max);
    a_min = fminx(a_min, 0.0f);
    a_max = fmaxx(a_max, 0.0f);
    return qo_tf_encoding(bw, a_min, a_max, sym, strict, unsigned_sym);
}

/* ---- MSE calibration: src/MseEncodingAnalyzer.cpp:77-285 (DTYPE = float), on tf_enhanced statistics (:70-76) ---- */
#define QO_MSE_MAX_EDGES 1024

static float mse_cost(int bw, const float* centers, const float* cpdf, int n_centers, float cand_min, float cand_max,
                      int sym, int strict, int unsigned_sym)
{
    qo_encoding enc = qo_tf_encoding(bw, cand_min, cand_max, sym, strict, unsigned_sym);
    float err       = 0;
    for (int i = 0; i < n_centers; i++)
    {
        float val     = centers[i];
        float clamped = fmaxx(cand_min, fminx(val, cand_max));
        int quantized = d2i(round(clamped / enc.delta - enc.offset));
        float deq     = (float) (enc.delta * (quantized + enc.offset));
        double diff   = (double) (val - deq);
        err           = (float) (err + cpdf[i] * (diff * diff)); /* float += float * pow(float, 2) */
    }
    return err;
}

qo_encoding qo_mse_compute(const qo_tfe_state* s, int bw_in, int sym, int strict, int unsigned_sym)
{
    qo_encoding e   = {0, 0, 0, 0, 0};
    uint8_t bw      = (uint8_t) bw_in;
    float num_steps = (float) (pow(2, bw) - 1);
    if (sym && strict)
        num_steps -= 1;
    if (!s->initialized)
    {
        if (s->stats_updated)
        {
            e.min    = -1;
            e.max    = 1;
            e.delta  = (e.max - e.min) / (int) num_steps;
            e.offset = floor(e.min / e.delta);
            e.min    = e.offset * e.delta;
            e.max    = e.min + (int) num_steps * e.delta;
            e.bw     = bw;
        }
        return e;
    }
    /* _minimizeMSE :139-201 */
    const float width = (float) (s->x_left[1] - s->x_left[0]);
    float hist_min    = (float) s->x_left[0];
    float hist_max    = (float) (s->x_left[QO_PDF_SIZE - 1] + width);
    float min_val, max_val;
    tfe_range(s, &min_val, &max_val);
    max_val = max_val + width;

    static __thread float edges[QO_MSE_MAX_EDGES + 2], centers[QO_MSE_MAX_EDGES + 2], cpdf[QO_MSE_MAX_EDGES + 2];
    static __thread float mins[QO_MSE_MAX_EDGES + 2], maxs[QO_MSE_MAX_EDGES + 2];
    int n_edges      = 0;
    edges[n_edges++] = min_val;
    int guard        = 0;
    for (float i = hist_min; i <= hist_max && n_edges < QO_MSE_MAX_EDGES && guard < 4 * QO_MSE_MAX_EDGES;
         i += width, guard++)
        if (i >= min_val && i <= max_val)
            edges[n_edges++] = i;

    /* _pickMinMaxCandidatesMSECalib :204-238 */
    int n_min = 0, n_max = 0;
    for (int k = 0; k < n_edges; k++)
    {
        if (edges[k] < 0)
            mins[n_min++] = edges[k];
        else if (edges[k] > 0)
            maxs[n_max++] = edges[k];
    }
    mins[n_min++] = 0;
    maxs[n_max++] = 0;

    float pdf_start = (float) s->x_left[0];
    float pdf_step  = (float) (s->x_left[1] - s->x_left[0]);
    int n_centers   = n_edges - 1;
    for (int i = 0; i < n_centers; i++)
    {
        centers[i] = (i == 0) ? min_val + width / 2 : centers[i - 1] + width;
        int ind    = f2i(floorf((centers[i] - pdf_start) / pdf_step));
        ind        = imin(imax(0, ind), QO_PDF_SIZE - 1);
        cpdf[i]    = (float) s->pdf[ind];
    }

    float mse_min  = FLT_MAX;
    float best_min = min_val, best_max = max_val;
    for (int a = 0; a < n_min; a++)
        for (int b = 0; b < n_max; b++)
        {
            if (a == n_min - 1 && b == n_max - 1)
                break; /* the trailing {0, 0} was popped */
            float c = mse_cost(bw, centers, cpdf, n_centers, mins[a], maxs[b], sym, strict, unsigned_sym);
            if (c < mse_min)
            {
                mse_min  = c;
                best_min = mins[a];
                best_max = maxs[b];
            }
        }
    best_min = fminx(best_min, 0.0f);
    best_max = fmaxx(best_max, 0.0f);
    return qo_tf_encoding(bw, best_min, best_max, sym, strict, unsigned_sym);
}

/* ---- broadcast QDQ: src/trim_functions.cpp:633-662 ---- */
void qo_qdq_broadcast(const float* in, float* out, int64_t num_element, int64_t num_dims, const int64_t* input_strides,
                      const int64_t* encoding_strides, const float* enc_min, const float* enc_max, const float* enc_delta,
                      const float* enc_offset)
{
    for (size_t i = 0; i < (size_t) num_element; i++)
    {
        int enc_idx   = 0;
        int remainder = (int) i;
        for (int64_t dim = 0; dim < num_dims; dim++)
        {
            int dim_idx = (int) (remainder / input_strides[dim]);
            remainder   = (int) (remainder - dim_idx * input_strides[dim]);
            enc_idx += (int) (encoding_strides[dim] * dim_idx);
        }
        qo_encoding e = {enc_min[enc_idx], enc_max[enc_idx], enc_delta[enc_idx], enc_offset[enc_idx], 0};
        qo_qdq(in + i, 1, out + i, &e);
    }
}


/* =====================================================================================================================
 * Entropy scheme (src/EntropyEncodingAnalyzer.cpp). DTYPE = float throughout, as libpymo instantiates it.
 * ===================================================================================================================== */
void qo_entropy_init(qo_entropy_state* s)
{
    memset(s, 0, sizeof(*s));
}

/* getBin, math_functions.cpp:466-470: every argument is narrowed to float; the quotient goes through a float -> size_t cast,
 * written here exactly as there so that the compiler emits the same conversion */
static size_t entropy_get_bin(size_t n_bins, float bin_width, float min_value, float value)
{
    size_t q;
    if (bin_width == 0)
        return 0;
    q = (size_t) ((value - min_value) / bin_width);
    return q < n_bins - 1 ? q : n_bins - 1;
}

/* updateTensorHistogram_cpu, math_functions.cpp:472-560 */
void qo_entropy_update(qo_entropy_state* s, const float* data, size_t n)
{
    double min_input, max_input;
    float bin_width;
    size_t i;
    s->stats_updated = 1;                      /* EntropyEncodingAnalyzer.cpp:84 */
    min_input        = qo_get_min(data, n);
    max_input        = qo_get_max(data, n);
    if (min_input == 0 && max_input == 0)      /* :478-483 */
        return;
    if (min_input == max_input)                /* :486-489: std::max(maxInput, minInput + (float) 0.01) */
    {
        double cand = min_input + (float) 0.01;
        max_input   = max_input < cand ? cand : max_input;
    }
    if (!s->initialized)                       /* :492-497 */
    {
        for (i = 0; i < QO_PDF_SIZE; ++i)
            s->histogram[i] = 0;
        s->min = min_input, s->max = max_input;
        s->initialized = 1;
    }
    if (min_input < s->min || max_input > s->max)   /* :500-548 */
    {
        double new_min  = s->min < min_input ? s->min : min_input;   /* std::min(minInput, tpp.min) */
        double new_max  = max_input < s->max ? s->max : max_input;   /* std::max(maxInput, tpp.max) */
        double dest_w   = (new_max - new_min) / QO_PDF_SIZE;
        double src_w    = (s->max - s->min) / QO_PDF_SIZE;
        double scaled[QO_PDF_SIZE];
        for (i = 0; i < QO_PDF_SIZE; ++i)
            scaled[i] = 0;
        for (i = 0; i < QO_PDF_SIZE; ++i)
        {
            double src_begin, dest_end, cnt, r;
            size_t dest_bin, b;
            if (s->histogram[i] == 0)
                continue;
            src_begin = s->min + src_w * i;
            dest_bin  = (size_t) ((src_begin - new_min) / dest_w);
            dest_end  = new_min + dest_w * (dest_bin + 1);
            r         = round((dest_end - src_begin) / src_w * s->histogram[i]);
            cnt       = s->histogram[i] < r ? s->histogram[i] : r;   /* std::min(round(..), histogram[i]) */
            b         = entropy_get_bin(QO_PDF_SIZE, dest_w, new_min, src_begin);
            scaled[b] += cnt;
            if (cnt < s->histogram[i])
            {
                b = entropy_get_bin(QO_PDF_SIZE, dest_w, new_min, src_begin + dest_w);
                scaled[b] += s->histogram[i] - cnt;
            }
        }
        for (i = 0; i < QO_PDF_SIZE; ++i)
            s->histogram[i] = scaled[i];
        s->min = new_min, s->max = new_max;
    }
    bin_width = (s->max - s->min) / QO_PDF_SIZE;   /* :550 */
    for (i = 0; i < n; ++i)
        s->histogram[entropy_get_bin(QO_PDF_SIZE, bin_width, s->min, data[i])] += 1;
    s->iterations++;
}

/* rescaleHistogram, math_functions.cpp:562-640 */
void qo_rescale_histogram(const double* src, double src_min, double src_max, double dst_min, double dst_max, double* dst)
{
    const size_t n = QO_PDF_SIZE;
    double src_w, dest_w;
    size_t si, d;
    if (src_min == dst_min && src_max == dst_max)
    {
        memcpy(dst, src, n * sizeof(double));
        return;
    }
    src_w  = (src_max - src_min) / n;
    dest_w = (dst_max - dst_min) / n;
    for (d = 0; d < n; ++d)
        dst[d] = 0;
    for (si = 0; si < n; ++si)
    {
        double val = src[si], s_start, s_stop, f0, f1, rem;
        size_t d0, d1;
        if (val == 0)
            continue;
        s_start = src_min + si * src_w;
        s_stop  = src_min + (si + 1) * src_w;
        f0      = floor((s_start - dst_min) / dest_w);
        f1      = ceil((s_stop - dst_min) / dest_w);
        d0      = (size_t) (f0 < 0.0 ? 0.0 : f0);   /* std::max(f, 0.0) */
        d1      = (size_t) (f1 < 0.0 ? 0.0 : f1);
        if (d0 >= n)
            d0 = n - 1;
        if (d1 >= n)
            d1 = n - 1;
        rem = val;
        for (d = d0; d <= d1; ++d)
        {
            double d_start = dst_min + d * dest_w;
            double d_stop  = dst_min + (d + 1) * dest_w;
            double o_start = s_start < d_start ? d_start : s_start;
            double o_stop  = d_stop < s_stop ? d_stop : s_stop;
            double ratio   = (o_stop - o_start) / src_w;
            double dist;
            ratio = ratio >= 0.0f ? ratio : 0.0f;
            ratio = ratio <= 1.0f ? ratio : 1.0f;
            dist  = round(ratio * val);
            dist  = dist <= rem ? dist : rem;
            dst[d] += dist;
            rem -= dist;
        }
    }
}

/* std::accumulate(first, last, 0.f): the accumulator has the type of the initial value, float */
static double entropy_accumulate(const double* p, size_t n)
{
    float acc = 0.f;
    size_t i;
    for (i = 0; i < n; ++i)
        acc = acc + p[i];
    return acc;
}

/* _conditionHistogram, EntropyEncodingAnalyzer.cpp:151-194 */
static void entropy_condition(double* hist, size_t length)
{
    const double eps_zero = 0.0001;
    size_t zeros = 0, i;
    int is_zero[QO_PDF_SIZE];
    double eps_non;
    if (length == 0)
        return;
    for (i = 0; i < length; ++i)
    {
        is_zero[i] = hist[i] == 0.f;
        zeros += is_zero[i];
    }
    if (zeros == length)
        return;
    eps_non = eps_zero * (double) zeros / (double) (length - zeros);
    if (eps_non >= 1.0)
        return;
    for (i = 0; i < length; ++i)
    {
        hist[i] += eps_zero * is_zero[i];
        hist[i] -= eps_non * (1 - is_zero[i]);
    }
}

/* _computeKL, :196-219 */
static double entropy_kl(double* P, double* Q, size_t length)
{
    double sum_p = entropy_accumulate(P, length), sum_q = entropy_accumulate(Q, length), divergence = 0;
    size_t i;
    for (i = 0; i < length; ++i)
    {
        P[i] /= sum_p;
        Q[i] /= sum_q;
        if (P[i] > 0 && Q[i] > 0)
            divergence += P[i] * log(P[i] / Q[i]);
    }
    return divergence;
}

/* _optimizeKL, :221-428 */
static void entropy_optimize_kl(const qo_entropy_state* s, int bw, int sym, int strict, int unsigned_sym, float* o_min,
                                float* o_max)
{
    double hist_min = s->min, hist_max = s->max, hist[QO_PDF_SIZE], P[QO_PDF_SIZE], Q[QO_PDF_SIZE];
    const size_t num_bins = QO_PDF_SIZE, num_q = 255;
    double bin_w, best = INFINITY, t_min, t_max;
    size_t start = 0, stop = num_bins - 1, i;
    if (sym && (hist_min < 0.0 || !unsigned_sym))   /* :229-241 */
    {
        double a = fabs(hist_max), b = fabs(hist_min);
        float abs_max = a < b ? b : a;   /* std::max(std::abs(histMax), std::abs(histMin)) narrowed to DTYPE */
        float abs_min = -abs_max;
        qo_rescale_histogram(s->histogram, hist_min, hist_max, abs_min, abs_max, hist);
        hist_min = abs_min, hist_max = abs_max;
    }
    else
        memcpy(hist, s->histogram, sizeof(hist));
    if (bw != 8)   /* :251-254 (the bin-count conditions cannot hold for 512 bins) */
    {
        *o_min = hist_min, *o_max = hist_max;
        return;
    }
    bin_w = (hist_max - hist_min) / (double) num_bins;
    t_min = hist_min, t_max = hist_max;
    while ((stop - start + 1) >= num_q)
    {
        const size_t win = stop - start + 1;
        const double* wp = hist + start;
        double left = 0, right = 0, merged, sum_p, sum_q, divergence;
        size_t q;
        for (i = 0; i < win; ++i)
            P[i] = 0, Q[i] = 0;
        for (i = 0; i <= start; ++i)
            left += hist[i];
        P[0] += left;
        for (i = start + 1; i < stop; ++i)
            P[i - start] = hist[i];
        for (i = stop; i < num_bins; ++i)
            right += hist[i];
        P[win - 1] += right;
        merged = (double) win / (double) num_q;
        for (q = 0; q < num_q; ++q)
        {
            const size_t i0 = ceil(q * merged);
            const size_t i1 = (q < num_q - 1) ? (size_t) ceil((q + 1) * merged) : win;
            double sum = 0, norm = 0;
            for (i = i0; i < i1; ++i)
            {
                sum += wp[i];
                norm += (wp[i] != 0);
            }
            if (norm != 0)
                for (i = i0; i < i1; ++i)
                    if (wp[i])
                        Q[i] = sum / norm;
        }
        sum_p = entropy_accumulate(P, win);
        sum_q = entropy_accumulate(Q, win);
        if (sum_p == 0 || sum_q == 0)
            break;
        entropy_condition(P, win);
        entropy_condition(Q, win);
        divergence = entropy_kl(P, Q, win);
        if (divergence < best)
        {
            best  = divergence;
            t_min = hist_min + start * bin_w;
            t_max = hist_min + (stop + 1) * bin_w;
        }
        if (sym || strict)
        {
            start++;
            stop--;
        }
        else
        {
            double loss[3];
            int k = 0, j;
            loss[0] = hist[start] + hist[stop];
            loss[1] = hist[start] + hist[start + 1];
            loss[2] = hist[stop] + hist[stop - 1];
            for (j = 1; j < 3; ++j)   /* std::min_element: first minimum */
                if (loss[j] < loss[k])
                    k = j;
            if ((k == 0 && (hist_min + (start + 1) * bin_w) > 0) || (k == 1 && (hist_min + (start + 2) * bin_w) > 0))
                k = 2;
            else if ((k == 0 && (hist_min + stop * bin_w) < 0) || (k == 2 && (hist_min + (stop - 1) * bin_w) < 0))
                k = 1;
            if (k == 0)
                start++, stop--;
            else if (k == 1)
                start += 2;
            else
                stop -= 2;
        }
    }
    *o_min = t_min, *o_max = t_max;
}

/* computeEncoding, :98-143 */
qo_encoding qo_entropy_compute(const qo_entropy_state* s, int bw, int sym, int strict, int unsigned_sym)
{
    qo_encoding e = {0, 0, 0, 0, 0};
    float num_steps = pow(2, bw) - 1, a_min, a_max;
    if (sym && strict)
        num_steps -= 1;
    if (!s->initialized)
    {
        if (s->stats_updated)
        {
            e.min    = -1;
            e.max    = 1;
            e.delta  = (e.max - e.min) / (int) num_steps;
            e.offset = floor(e.min / e.delta);
            e.min    = e.offset * e.delta;
            e.max    = e.min + (int) num_steps * e.delta;
            e.bw     = bw;
        }
        return e;
    }
    entropy_optimize_kl(s, bw, sym, strict, unsigned_sym, &a_min, &a_max);
    a_min = (0.f < a_min) ? 0.f : a_min;   /* std::min(aMin, 0.f) */
    a_max = (a_max < 0.f) ? 0.f : a_max;   /* std::max(aMax, 0.f) */
    return qo_tf_encoding(bw, a_min, a_max, sym, strict, unsigned_sym);
}
