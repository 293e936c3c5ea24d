// libat_b200: the token consumer's batch assembly, device resident.
//
// Replaces, for tokens that are already in HBM (the tokenizer's output, or token files read in bulk), what the reference
// does per training batch on the host (datasets/tokenized_spec_dataset.py:52-76 + datasets/data_loader_creator.py:17-34):
//   __getitem__   np.load of one clip's token file -> seq; multi-hot label vector of num_classes floats
//   collate_fn    pad_sequence(sequences, batch_first=True, padding_value=0).long(); attention masks; torch.stack(labels)
// One launch gathers a batch of variable-length token sequences out of the flat token store into the padded (B, T_max)
// int64 matrix, writes the attention masks and the multi-hot labels.
#include "at_common.cuh"

namespace at {

// grid (B), block 256.  tokens: flat store; offsets[n_clips + 1]; idx[B]: clip ids of the batch; out (B, t_max) int64;
// mask (B, t_max) float or null: 1 where a token exists, 0 in the padding -- or all ones when mask_all_ones (the reference's
// collate_fn builds its masks from the already padded matrix, tokenized_spec_dataset.py:70-74, so they are all ones).
template <typename TOK>
__global__ void __launch_bounds__(256) k_tokens_collate(const TOK *__restrict__ tokens, const int64_t *__restrict__ offsets,
                                                        const int64_t *__restrict__ idx, int t_max, int mask_all_ones,
                                                        int64_t *__restrict__ out, float *__restrict__ mask) {
    const int b = blockIdx.x;
    const int64_t clip = idx[b];
    const int64_t o0 = offsets[clip];
    int64_t len = offsets[clip + 1] - o0;
    if (len > t_max) len = t_max;
    int64_t *orow = out + (int64_t)b * t_max;
    float *mrow = mask ? mask + (int64_t)b * t_max : nullptr;
    for (int t = threadIdx.x; t < t_max; t += blockDim.x) {
        const bool live = t < len;
        orow[t] = live ? (int64_t)tokens[o0 + t] : 0;
        if (mrow) mrow[t] = (live || mask_all_ones) ? 1.0f : 0.0f;
    }
}

// labels[b, :] = 0; labels[b, label_ids[j]] = 1 for j in [label_offsets[clip], label_offsets[clip + 1]).  grid (B), block 128.
__global__ void __launch_bounds__(128) k_multihot(const int32_t *__restrict__ label_ids, const int64_t *__restrict__ label_offsets,
                                                  const int64_t *__restrict__ idx, int num_classes, float *__restrict__ labels) {
    const int b = blockIdx.x;
    float *row = labels + (int64_t)b * num_classes;
    for (int c = threadIdx.x; c < num_classes; c += blockDim.x) row[c] = 0.f;
    __syncthreads();
    const int64_t clip = idx[b];
    for (int64_t j = label_offsets[clip] + threadIdx.x; j < label_offsets[clip + 1]; j += blockDim.x) {
        const int c = label_ids[j];
        if (c >= 0 && c < num_classes) row[c] = 1.0f;
    }
}

// max over the batch of the clip lengths (device scalar), so a caller can size the padded matrix without a host copy of
// the offsets
__global__ void k_batch_max_len(const int64_t *__restrict__ offsets, const int64_t *__restrict__ idx, int B, int *__restrict__ out) {
    int m = 0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const int64_t clip = idx[b];
        const int64_t len = offsets[clip + 1] - offsets[clip];
        m = max(m, (int)(len > 0x7FFFFFFF ? 0x7FFFFFFF : len));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

}  // namespace at

using namespace at;

extern "C" {

int at_tokens_collate(const void *tokens, int token_bytes, const int64_t *offsets, const int64_t *idx, int B, int t_max,
                      int mask_all_ones, int64_t *out, float *mask, void *stream) {
    AT_REQUIRE(tokens && offsets && idx && out && B >= 0 && t_max >= 0, "at_tokens_collate: bad arguments");
    AT_REQUIRE(token_bytes == 4 || token_bytes == 8, "at_tokens_collate: tokens must be int32 or int64");
    if (B == 0 || t_max == 0) return AT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (token_bytes == 8)
        k_tokens_collate<int64_t><<<B, 256, 0, st>>>((const int64_t *)tokens, offsets, idx, t_max, mask_all_ones, out, mask);
    else
        k_tokens_collate<int32_t><<<B, 256, 0, st>>>((const int32_t *)tokens, offsets, idx, t_max, mask_all_ones, out, mask);
    AT_LAUNCH_OK();
    return AT_OK;
}

int at_tokens_multihot(const int32_t *label_ids, const int64_t *label_offsets, const int64_t *idx, int B, int num_classes,
                       float *labels, void *stream) {
    AT_REQUIRE(label_ids && label_offsets && idx && labels && B >= 0 && num_classes > 0, "at_tokens_multihot: bad arguments");
    if (B == 0) return AT_OK;
    k_multihot<<<B, 128, 0, (cudaStream_t)stream>>>(label_ids, label_offsets, idx, num_classes, labels);
    AT_LAUNCH_OK();
    return AT_OK;
}

int at_tokens_batch_max_len(const int64_t *offsets, const int64_t *idx, int B, int32_t *out_dev, void *stream) {
    AT_REQUIRE(offsets && idx && out_dev && B >= 0, "at_tokens_batch_max_len: bad arguments");
    AT_CUDA_OK(cudaMemsetAsync(out_dev, 0, sizeof(int32_t), (cudaStream_t)stream));
    if (B == 0) return AT_OK;
    k_batch_max_len<<<1, 256, 0, (cudaStream_t)stream>>>(offsets, idx, B, out_dev);
    AT_LAUNCH_OK();
    return AT_OK;
}

}  // extern "C"
