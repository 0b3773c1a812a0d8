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
