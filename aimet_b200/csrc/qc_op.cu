// The compute body of the reference's ONNX Runtime custom op "QcQuantizeOp" on the device, behind the C ABI.
//
// Reference: TrainingExtensions/onnx/src/QcQuantizeOp.cpp:64-143 (computeImpl and the CUDA Compute wrapper) and
// TrainingExtensions/onnx/src/AimetOpUtils.h:98-322 (modeSpecificActionInt / PerChannelInt / BroadcastInt),
// QuantizeDequantizeUtils.{hpp,cpp} (quantizeDequantizePerChannel :110-166, BroadcastShapeInfo, quantizeDequantizeBroadcast).
// onnxruntime is not part of this image, so the ORT glue itself (Ort::Custom::CreateLiteCustomOp, the kernel-info attribute
// carrying a QcQuantizeInfo pointer) is what a maintainer adds around this call -- INTEGRATION.md shows it. What is here is
// everything QcQuantizeOpCuda::Compute does between "input buffer" and "output buffer":
//   passThrough / disabled          output = input
//   updateStats                     statistics of the tensor (per tensor, per channel along axis 0, or per contiguous block),
//                                   device resident, NO stream synchronisation (the reference has to cudaStreamSynchronize
//                                   first: its statistics ignore the stream, QcQuantizeOp.cpp:136-141); output = input
//   quantizeDequantize              per tensor (min / max / bw of the encoding), per channel along ANY axis, blockwise
//                                   (encoding tensor broadcast over the input)
//   oneShotQuantizeDequantize       reset + statistics + encoding search + QDQ; the HOST encoding objects are updated (one
//                                   device->host copy, as the reference's computeEncoding returns host values)
// Scratch is caller-owned (`workspace`): nothing is allocated here, where the reference cudaMallocs per call.
#include <vector>

#include "common.cuh"

using namespace ab;

namespace
{
int64_t num_elements(const int64_t* shape, int ndim)
{
    int64_t n = 1;
    for (int i = 0; i < ndim; ++i)
        n *= shape[i];
    return n;
}

int copy_through(const void* in, void* out, int64_t count, int dtype, cudaStream_t st)
{
    if (in != out && count > 0)
        AB_CUDA_CHECK(cudaMemcpyAsync(out, in, (size_t) count * (dtype == AB_F32 ? 4 : 2), cudaMemcpyDeviceToDevice, st));
    return AB_OK;
}

// {min | max | delta | offset} of the host encodings as floats, staged in pinned-or-pageable host memory and copied to the
// workspace (QuantizeDequantizeUtils.hpp:129-157 narrows the doubles to T the same way)
int upload_params(const ab_encoding* enc, int n, float* dev, std::vector<float>& host, cudaStream_t st)
{
    host.resize(4 * (size_t) n);
    for (int i = 0; i < n; ++i)
    {
        host[i]         = (float) enc[i].min;
        host[n + i]     = (float) enc[i].max;
        host[2 * n + i] = (float) enc[i].delta;
        host[3 * n + i] = (float) enc[i].offset;
    }
    AB_CUDA_CHECK(cudaMemcpyAsync(dev, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    AB_CUDA_CHECK(cudaStreamSynchronize(st));   // `host` is pageable and dies with this call
    return AB_OK;
}

// search on the device, encodings back to the host objects (min / max / delta / offset; bw stays: AimetOpUtils.h:115-118)
int search_to_host(ab_qc_quantize_info* info, int n, double* enc_dev, cudaStream_t st)
{
    const int bw = info->encodings[0].bw;
    int rc = ab_compute_encodings(info->states, n, info->quant_mode, bw, info->use_symmetric_encoding, 0, 0, enc_dev, nullptr, st);
    if (rc != AB_OK)
        return rc;
    std::vector<double> host(5 * (size_t) n);
    AB_CUDA_CHECK(cudaMemcpyAsync(host.data(), enc_dev, host.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    AB_CUDA_CHECK(cudaStreamSynchronize(st));
    for (int i = 0; i < n; ++i)
    {
        info->encodings[i].min    = host[5 * i];
        info->encodings[i].max    = host[5 * i + 1];
        info->encodings[i].delta  = host[5 * i + 2];
        info->encodings[i].offset = host[5 * i + 3];
    }
    return AB_OK;
}
}   // namespace

extern "C" size_t ab_qc_quantize_op_workspace_bytes(int num_encodings)
{
    // encoding rows of the search (5 doubles each) + the float[4][n] parameter block
    return (size_t) (num_encodings > 0 ? num_encodings : 1) * (5 * sizeof(double) + 4 * sizeof(float)) + 64;
}

extern "C" int ab_qc_quantize_op_compute(ab_qc_quantize_info* info, const void* in, void* out, const int64_t* shape, int ndim,
                                         int dtype, void* workspace, void* stream)
{
    if (info == nullptr || shape == nullptr || ndim < 0 || ndim > 7)
    {
        set_error("null info / shape, or more than 7 dimensions");
        return AB_ERR_INVALID;
    }
    if (dtype != AB_F32 && dtype != AB_BF16)
    {
        set_error("unsupported dtype %d", dtype);
        return AB_ERR_UNSUPPORTED;
    }
    cudaStream_t st     = (cudaStream_t) stream;
    const int64_t count = num_elements(shape, ndim);
    if (count > 0 && (in == nullptr || out == nullptr))
    {
        set_error("null tensor pointer");
        return AB_ERR_INVALID;
    }
    int mode = info->op_mode;
    if (!info->enabled)
        mode = AB_OP_PASS_THROUGH;                       // QcQuantizeOp.cpp:77-81
    if (mode == AB_OP_PASS_THROUGH || count == 0)
        return copy_through(in, out, count, dtype, st);
    if (mode < AB_OP_UPDATE_STATS || mode > AB_OP_PASS_THROUGH)
    {
        set_error("unknown op mode %d", mode);
        return AB_ERR_INVALID;
    }
    if (!info->is_int_data_type)
    {
        set_error("float (fp16) simulation is outside the aimet_b200 hot path");
        return AB_ERR_UNSUPPORTED;
    }
    if (info->encodings == nullptr || info->num_encodings < 1 ||
        ((mode == AB_OP_UPDATE_STATS || mode == AB_OP_ONE_SHOT_QDQ) && info->states == nullptr) ||
        (workspace == nullptr && (mode == AB_OP_ONE_SHOT_QDQ || info->use_per_channel_mode)))
    {
        set_error("encodings, statistics records or workspace missing for this mode");
        return AB_ERR_INVALID;
    }
    double* enc_dev   = reinterpret_cast<double*>(workspace);
    float* params_dev = workspace ? reinterpret_cast<float*>(enc_dev + 5 * (size_t) info->num_encodings) : nullptr;
    std::vector<float> host_params;
    int rc;

    // ---- per tensor (AimetOpUtils.h:98-143) ------------------------------------------------------------------------------
    if (!info->use_per_channel_mode)
    {
        ab_encoding* e = info->encodings;
        switch (mode)
        {
        case AB_OP_UPDATE_STATS:
            rc = ab_stats_update(in, count, dtype, info->quant_mode, info->states, nullptr, 0, stream);
            return rc != AB_OK ? rc : copy_through(in, out, count, dtype, st);
        case AB_OP_ONE_SHOT_QDQ:
            if ((rc = ab_stats_reset(info->states, 1, stream)) != AB_OK ||
                (rc = ab_stats_update(in, count, dtype, info->quant_mode, info->states, nullptr, 0, stream)) != AB_OK ||
                (rc = search_to_host(info, 1, enc_dev, st)) != AB_OK)
                return rc;
            info->op_mode = AB_OP_QDQ;                   // one shot only once (QcQuantizeOp.cpp:109-113)
            /* fall through */
        default:
            return ab_qdq_per_tensor_fwd(in, out, count, dtype, e->min, e->max, e->bw, AB_ROUND_NEAREST, 0, stream);
        }
    }

    // ---- blockwise: an encoding per (channel, block) broadcast over the input (AimetOpUtils.h:218-300) --------------------
    if (info->block_size > 0)
    {
        const int ca = info->channel_axis, ba = info->block_axis;
        if (ca < 0 || ca >= ndim || ba < 0 || ba >= ndim || ca == ba || shape[ba] % info->block_size != 0)
        {
            set_error("Block dimension is not evenly divisible by block size (or bad axes).");
            return AB_ERR_INVALID;
        }
        // BroadcastShapeInfo (QuantizeDequantizeUtils.cpp:100-150): the block axis splits into (blocks, block size)
        int64_t tshape[8], eshape[8], tstride[8], estride[8];
        int nd = 0;
        for (int i = 0; i < ndim; ++i)
        {
            if (i == ca)
                tshape[nd] = shape[i], eshape[nd++] = shape[i];
            else if (i == ba)
            {
                tshape[nd] = shape[i] / info->block_size, eshape[nd++] = shape[i] / info->block_size;
                tshape[nd] = info->block_size, eshape[nd++] = 1;
            }
            else
                tshape[nd] = shape[i], eshape[nd++] = 1;
        }
        int64_t ts = 1, es = 1, n_enc = 1;
        for (int i = nd - 1; i >= 0; --i)
        {
            tstride[i] = ts, ts *= tshape[i];
            estride[i] = (eshape[i] == 1 && tshape[i] != 1) ? 0 : es, es *= eshape[i];
            n_enc *= eshape[i];
        }
        if (n_enc != info->num_encodings)
        {
            set_error("Expected number of encodings (%lld) does not match provided encoding list size (%d).", (long long) n_enc,
                      info->num_encodings);
            return AB_ERR_INVALID;
        }
        if (mode == AB_OP_UPDATE_STATS || mode == AB_OP_ONE_SHOT_QDQ)
        {
            // blocks are contiguous when everything after the block axis (and the channel axis, if it comes later) has extent 1
            bool contiguous = true;
            for (int i = 0; i < nd; ++i)
                if (estride[i] == 0 && i + 1 < nd)
                    for (int j = i + 1; j < nd; ++j)
                        if (eshape[j] != 1)
                            contiguous = false;
            if (!contiguous)
            {
                set_error("statistics of non-contiguous blocks need the permuted copy the reference makes "
                          "(copyToContiguousBlockLayout): not on this path");
                return AB_ERR_UNSUPPORTED;
            }
            if (mode == AB_OP_ONE_SHOT_QDQ && (rc = ab_stats_reset(info->states, n_enc, stream)) != AB_OK)
                return rc;
            if ((rc = ab_stats_update_segmented(in, n_enc, count / n_enc, dtype, info->quant_mode, info->states, stream)) != AB_OK)
                return rc;
            if (mode == AB_OP_UPDATE_STATS)
                return copy_through(in, out, count, dtype, st);
            if ((rc = search_to_host(info, (int) n_enc, enc_dev, st)) != AB_OK)
                return rc;
            info->op_mode = AB_OP_QDQ;
        }
        if ((rc = upload_params(info->encodings, (int) n_enc, params_dev, host_params, st)) != AB_OK)
            return rc;
        return ab_qdq_broadcast_fwd(in, out, count, nd, tstride, estride, params_dev, params_dev + n_enc, params_dev + 2 * n_enc,
                                    params_dev + 3 * n_enc, dtype, stream);
    }

    // ---- per channel (AimetOpUtils.h:146-215) -----------------------------------------------------------------------------
    const int axis = info->channel_axis;
    if (axis < 0 || axis >= ndim || shape[axis] != info->num_encodings)
    {
        set_error("Channel dimensions do not match encoding vector size.");
        return AB_ERR_INVALID;
    }
    const int64_t channels = shape[axis];
    int64_t inner          = 1;
    for (int i = axis + 1; i < ndim; ++i)
        inner *= shape[i];
    if (mode == AB_OP_UPDATE_STATS || mode == AB_OP_ONE_SHOT_QDQ)
    {
        int64_t outer = 1;
        for (int i = 0; i < axis; ++i)
            outer *= shape[i];
        if (outer != 1)
        {
            set_error("per-channel statistics along an inner axis need the slice copies the reference makes "
                      "(sliceTensorAlongAxis): only channel axes with no outer extent are on this path");
            return AB_ERR_UNSUPPORTED;
        }
        if (mode == AB_OP_ONE_SHOT_QDQ && (rc = ab_stats_reset(info->states, channels, stream)) != AB_OK)
            return rc;
        if ((rc = ab_stats_update_segmented(in, channels, inner, dtype, info->quant_mode, info->states, stream)) != AB_OK)
            return rc;
        if (mode == AB_OP_UPDATE_STATS)
            return copy_through(in, out, count, dtype, st);
        if ((rc = search_to_host(info, (int) channels, enc_dev, st)) != AB_OK)
            return rc;
        info->op_mode = AB_OP_QDQ;
    }
    if ((rc = upload_params(info->encodings, (int) channels, params_dev, host_params, st)) != AB_OK)
        return rc;
    return ab_qdq_per_channel_fwd(in, out, channels, count, inner, dtype, params_dev, AB_ROUND_NEAREST, 0, stream);
}
