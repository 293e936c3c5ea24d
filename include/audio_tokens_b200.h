/*
 * audio_tokens_b200.h -- C ABI of the B200-native audio-tokens hot path.
 *
 * One shared library (libat_b200.so, built from audio-tokens_b200/csrc) replaces the three library
 * call sites of danavery/audio-tokens' preprocessing path.  Every entry point takes plain pointers and
 * sizes; there are no torch / numpy types here.  Unless a name ends in `_host`, all data pointers are
 * DEVICE pointers owned by the caller and every call is stream-ordered on the `stream` argument
 * (a cudaStream_t passed as void*; NULL = the legacy default stream).  Calls return 0 on success or a
 * negative at_status; at_last_error() gives the thread-local message.  The library owns only the opaque
 * plan / index / k-means objects, freed by *_destroy.
 *
 * There is no CPU fallback: without a CUDA device every compute entry point fails with AT_ERR_CUDA.
 *
 * Threading (the reference is single-threaded Python): a plan / index / k-means object may be used by
 * one thread at a time; different objects may be used from different threads.  The error message is thread-local; the
 * launch counter of at_kernel_launches() and the at_profile_* timers are process-wide and meant for one measuring thread.
 * Per-device kernel attributes are set on first use of each device, so a process may drive several GPUs.
 *
 * Reference interface each group replaces (paths relative to the reference repo):
 *
 *   at_mel_*     processors/spectrogram_generator.py:28-34 (MelSpectrogram + AmplitudeToDB construction),
 *                :123-126 generate_mel_spectrogram, :128-131 normalize_spectrogram, :133-146 check_for_nan_inf;
 *                i.e. torchaudio.transforms.MelSpectrogram / AmplitudeToDB on (1, L) tensors.
 *   at_row_l2norm  processors/cluster_creator.py:64-66 and processors/spec_tokenizer.py:106-109
 *                (normalize_vectors: v / (||v|| + 1e-10)).
 *   at_index_*   faiss.IndexFlatL2(d) / .add / .search(x, 1) as used in processors/spec_tokenizer.py:123-127,77
 *                and inside faiss.Kmeans.train (processors/cluster_creator.py:54-56).
 *   at_kmeans_*  faiss.Kmeans(d, k, niter=, ...).train(x[, init_centroids]) -> Clustering::train
 *                (processors/cluster_creator.py:42-58): one Lloyd iteration = accumulate (+ optional
 *                all-reduce by the caller) + finalize.
 *   at_synth_*   no reference counterpart: synthetic 16-bit-PCM clips for tests and benchmarks.
 */
#ifndef AUDIO_TOKENS_B200_H
#define AUDIO_TOKENS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AT_B200_VERSION 100 /* 0.1.0 */

typedef enum at_status {
    AT_OK = 0,
    AT_ERR_INVALID = -1,     /* bad argument (null pointer, unsupported n_fft / d, k > limits ...) */
    AT_ERR_CUDA = -2,        /* a CUDA runtime call failed (including: no device) */
    AT_ERR_UNSUPPORTED = -3, /* valid request the kernels do not cover (stated in the message) */
    AT_ERR_NOMEM = -4
} at_status;

int at_version(void);
const char *at_last_error(void);
/* sm_count / compute capability of the current device; AT_ERR_CUDA when there is none. */
int at_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* Number of kernels this library has launched so far in this process (monotonic). */
int64_t at_kernel_launches(void);
/* Optional kernel timing with CUDA events on the launching stream (bench.py's roofline figures).
 * at_profile_enable(1) starts recording an event pair around every launch of the kernels tagged below;
 * at_profile_summary synchronises the recorded events and returns the launch count and summed device
 * milliseconds of one tag, then forgets them. Tags: 0 search (distance-argmin), 1 mel, 2 k-means update
 * (counts, objective, scan, place, gather-sum), 3 k-means finalize. */
int at_profile_enable(int on);
int at_profile_summary(int tag, int64_t *launches, double *total_ms);

/* ------------------------------------------------------------------------------------------------
 * Stage 1: waveform -> mel dB spectrogram (+ optional per-clip min-max, + optional row L2 norm)
 * ---------------------------------------------------------------------------------------------- */
typedef struct at_mel_plan at_mel_plan;

/* MelSpectrogram(sample_rate, n_mels, n_fft, hop_length) + AmplitudeToDB() with torchaudio defaults:
 * periodic Hann (win_length = n_fft), center + reflect pad, power 2, HTK filterbank, f_min 0,
 * f_max sample_rate/2, norm None, 10*log10(max(x,1e-10)).  n_fft in {256, 512, 1024}; n_mels <= 256 (<= 128 for
 * n_fft 256).
 * normalize != 0 applies (s - min s) / (max s - min s) over each clip (normalize_spectrogram). */
int at_mel_plan_create(int sample_rate, int n_fft, int hop_length, int n_mels, int normalize, at_mel_plan **plan);
/* Optional: replace the built-in constants by the caller's (HOST pointers): window[n_fft] and the dense
 * filterbank fb[(n_fft/2+1) * n_mels] (row = frequency bin), e.g. the tensors torchaudio itself builds. */
int at_mel_plan_set_constants_host(at_mel_plan *plan, const float *window, const float *fb);
/* What at_mel_forward writes: AT_MEL_OUT_DB (default) = MelSpectrogram followed by AmplitudeToDB, the fused form of
 * generate_mel_spectrogram (processors/spectrogram_generator.py:123-126); AT_MEL_OUT_POWER = the mel power spectrogram,
 * i.e. torchaudio.transforms.MelSpectrogram alone (processors/spectrogram_generator.py:28-33,124), for callers that
 * keep AmplitudeToDB as a separate operator (at_amplitude_to_db).  The power output cannot be min-max normalised. */
enum { AT_MEL_OUT_DB = 0, AT_MEL_OUT_POWER = 1 };
int at_mel_plan_set_output(at_mel_plan *plan, int kind);
int at_mel_plan_destroy(at_mel_plan *plan);
/* 1 + n_samples / hop_length (center=True). */
int64_t at_mel_num_frames(const at_mel_plan *plan, int64_t n_samples);
/* Number of independent clip streams one launch works on (SMs x warp groups per CTA): a launch is balanced when its
 * clip count is a multiple of this (or much larger), which is how callers should size streaming chunks. */
int at_mel_work_groups(const at_mel_plan *plan);
/* While absmax_dev (device float, zeroed by the caller) is attached, every forward that writes the L2-normalised copy
 * raises it to the largest |element| written: the value at_kmeans_begin needs (normalize_vectors' output range,
 * processors/cluster_creator.py:52-56,64-66), without a separate at_absmax pass over the rows.  NULL detaches. */
int at_mel_plan_set_absmax_out(at_mel_plan *plan, float *absmax_dev);

/* B clips.  Clip b occupies wave[sample_offsets[b] .. sample_offsets[b+1]) and writes frames
 * out[frame_offsets[b] .. frame_offsets[b+1]) x n_mels, FRAME-MAJOR ([T][n_mels], the physical layout of
 * the reference's fortran_order (n_mels, T) .npy files).  sample_offsets / frame_offsets are DEVICE arrays
 * of B+1 int64; pass NULL for both to mean B uniform clips of `uniform_samples` samples each.
 * bad_flags[b] (may be NULL): 0 ok, 1 tile holds NaN/Inf (the reference drops the clip), 2 clip shorter than
 * n_fft/2+1 samples (torch's reflect pad raises).  out_l2 (may be NULL) receives the row-L2-normalised copy
 * x / (||x|| + 1e-10) that ClusterCreator / SpecTokenizer would compute next. */
int at_mel_forward(at_mel_plan *plan, const float *wave, const int64_t *sample_offsets,
                   const int64_t *frame_offsets, int64_t uniform_samples, int B, float *out,
                   float *out_l2, int32_t *bad_flags, void *stream);

/* Same with HOST buffers (pinned or pageable) and uniform clips: chunks of clips are copied H2D, transformed
 * and copied back D2H on two streams so copies overlap compute.  Blocks until the result is in out. */
int at_mel_forward_host(at_mel_plan *plan, const float *wave, int64_t uniform_samples, int B, float *out,
                        int32_t *bad_flags);

/* torchaudio.transforms.AmplitudeToDB.forward (processors/spectrogram_generator.py:34,125; torchaudio
 * functional.amplitude_to_DB with top_db=None): out = multiplier * log10(max(x, amin)) - multiplier * db_multiplier.
 * AmplitudeToDB() defaults: multiplier 10 (stype "power"), amin 1e-10, db_multiplier log10(max(amin, 1)) = 0.
 * Device pointers; in place (out == x) allowed. */
int at_amplitude_to_db(const float *x, int64_t n, float multiplier, float amin, float db_multiplier, float *out,
                       void *stream);

/* ------------------------------------------------------------------------------------------------
 * normalize_vectors
 * ---------------------------------------------------------------------------------------------- */
int at_row_l2norm(const float *x, int64_t n, int d, float *out, void *stream);

/* use_convolution branch of stages 2-3 (processors/cluster_creator.py:28-34,68-81; processors/spec_tokenizer.py:92-104):
 * nn.Conv1d(1, num_kernels, kernel_size, padding = kernel_size // 2) along the mel axis of every frame, output laid out like
 * conv_output.transpose(1, 2).reshape(-1, num_kernels * n_mels): out[i, m * num_kernels + c] = bias[c] + sum_t weight[c][t] *
 * x[i, m + t - kernel_size // 2].  x (n, n_mels), weight (num_kernels, kernel_size), bias (num_kernels), out (n, n_mels *
 * num_kernels): device fp32.  Odd kernel_size <= 15 (an even one changes the width), num_kernels <= 256.  The wide rows go
 * through at_row_l2norm, at_kmeans_* and at_index_search (d up to 1024; rows wider than 128 values must be normalised
 * beforehand and are searched by the exact fp32 tile kernel or, for d a multiple of 64 and labels only, by the
 * slice-accumulating tcgen05 kernel, which returns the exact kernel's labels). */
int at_conv_expand(const float *x, int64_t n, int n_mels, const float *weight, const float *bias, int num_kernels,
                   int kernel_size, float *out, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Stage 3 (and the inner search of stage 2): nearest centroid under FAISS's
 * |x|^2 + |c|^2 - 2<x,c> (clamped at 0), lowest index wins exact ties.
 * ---------------------------------------------------------------------------------------------- */
typedef struct at_index at_index;

enum { AT_ALGO_AUTO = 0, AT_ALGO_SIMT = 1, AT_ALGO_TENSOR = 2 };

int at_index_create(int d, at_index **index);
int at_index_destroy(at_index *index);
/* IndexFlatL2.reset() + add(k, centroids): copies the (k, d) fp32 centroids and prepares the operands
 * (norms; for the tcgen05 path the split-fp16 tiles). */
int at_index_set_centroids(at_index *index, const float *centroids, int k, void *stream);
int at_index_ntotal(const at_index *index);
/* Operand policy of the tcgen05 kernel: 0 auto (default: resident when the operand tiles fit), 1 always stream
 * centroid tiles through the shared-memory ring, 2 same as 0. */
int at_index_set_tc_mode(at_index *index, int mode);
/* Cumulative counters of the tcgen05 kernel (synchronises the device): out[0] = rows whose label was decided by the
 * fp32 re-check of the two best column groups, out[1] = rows scanned exactly over all centroids.  Every other row
 * was certified from the accumulator (at_assign_tc.cu, "Certification"). */
int at_index_tc_stats(at_index *index, uint64_t out[2]);
/* fp32 (k, d) centroids currently held (device pointer, valid until the next set / destroy). */
const float *at_index_centroids(const at_index *index);
/* search(x, 1).  l2norm_rows != 0 applies normalize_vectors to each row while loading.
 * Any of labels32 / labels64 / dist may be NULL.  algo: AT_ALGO_*; AUTO picks the tcgen05 kernel when
 * d == 64 and k >= 64 (or d a multiple of 64 up to 1024, k >= 64, pre-normalised rows, labels only and at least 8 M
 * row x centroid pairs: the slice-accumulating form), else the exact fp32 SIMT kernel.  Both return the argmin of the
 * same fp32 formula (the tensor path certifies each row from its accumulators or re-checks / re-scans it with that
 * formula) and, in dist, that formula's value for the returned label. */
int at_index_search(at_index *index, const float *x, int64_t n, int l2norm_rows, int algo,
                    int32_t *labels32, int64_t *labels64, float *dist, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Stage 2: k-means Lloyd iterations (faiss::Clustering semantics)
 * ---------------------------------------------------------------------------------------------- */
typedef struct at_kmeans at_kmeans;

int at_kmeans_create(int d, int k, at_kmeans **km);
int at_kmeans_destroy(at_kmeans *km);
/* Set / read the current (k, d) fp32 centroids (device pointers). */
int at_kmeans_set_centroids(at_kmeans *km, const float *centroids, void *stream);
const float *at_kmeans_centroids(const at_kmeans *km);
/* Copy the current (k, d) centroids into a caller-owned device buffer. */
int at_kmeans_get_centroids(const at_kmeans *km, float *out, void *stream);
/* Must be called once per training set before accumulate: fixes the fixed-point scale of the exact
 * (order-independent) sums from max |x| and the global row count. max_abs: HOST float, the max |x_ij| over
 * ALL ranks' rows (at_absmax computes the local one). */
int at_absmax(const float *x, int64_t n_elems, float *out_dev, void *stream);
int at_kmeans_begin(at_kmeans *km, float max_abs, int64_t n_total);
/* Same, ordered on `stream` only (at_kmeans_begin synchronises the whole device before and after): lets copies / kernels
 * of the next training set stay in flight on other streams. */
int at_kmeans_begin_on(at_kmeans *km, float max_abs, int64_t n_total, void *stream);
/* Number of int64 words in the accumulator buffer: k*d sums + k counts + 1 (the sum of |x_i|^2, from which finalize
 * derives the objective). */
int64_t at_kmeans_accum_words(const at_kmeans *km);
/* Search the local rows against the current centroids and accumulate per-cluster fixed-point sums, counts
 * and the sum of |x_i|^2 into accum (device int64[at_kmeans_accum_words], overwritten).  labels32 may be NULL.
 * Sums are exact integers, so adding the buffers of several ranks (all-reduce SUM on int64) gives a result
 * that does not depend on the rank count or on summation order. */
int at_kmeans_accumulate(at_kmeans *km, const float *x, int64_t n_local, int l2norm_rows, int algo,
                         int64_t *accum, int32_t *labels32, void *stream);
/* compute_centroids' division + split_clusters (mt19937(1234), EPS = 1/1024) + index refresh from the
 * (already all-reduced) accum.  stats (device float[4], may be NULL): objective, nsplit, imbalance factor,
 * number of empty clusters before the split. */
int at_kmeans_finalize(at_kmeans *km, const int64_t *accum, int64_t n_total, float *stats, void *stream);
/* Between two accumulate calls over the SAME rows (same pointer, same n_local, no at_kmeans_begin /
 * at_kmeans_set_centroids in between) only the rows whose label changed are moved: their value is subtracted from
 * the old cluster's exact integer sum and added to the new one, which gives bit-identical sums to regrouping every
 * row.  on = 0 switches this off (every accumulate regroups every row); default on. */
int at_kmeans_set_incremental(at_kmeans *km, int on);
/* Tokenizing the rows that were just clustered (run_pipeline.py:12-13: ClusterCreator, then SpecTokenizer over the same
 * spectrograms): search(x, 1) against `index`, re-using the fp16 operand image that at_kmeans_accumulate built for km's
 * current training set instead of converting the rows a second time.  x must be those n rows -- either the very array km
 * was trained on (l2norm_rows = 0), or the un-normalised rows it was derived from with normalize_vectors (l2norm_rows != 0:
 * the candidate re-checks and exact scans normalise x canonically; a last-bit difference between the two normalisations is
 * inside the certification threshold).  Same results as at_index_search(..., AT_ALGO_TENSOR, ...).  d == 64 only. */
int at_index_search_trained_rows(at_index *index, at_kmeans *km, const float *x, int64_t n, int l2norm_rows,
                                 int32_t *labels32, int64_t *labels64, float *dist, void *stream);
/* CONTRACT of the two caches behind at_kmeans_accumulate (the fp16 operand image of the rows, and the incremental
 * update's previous labels and local sums): they are keyed on (x pointer, n_local) only, so the CONTENTS of x must not
 * change between at_kmeans_begin and the last accumulate of that training set.  A caller that re-uses a buffer for new
 * rows (or mutates it in place) without calling at_kmeans_begin again must call at_kmeans_invalidate first: the next
 * accumulate then rebuilds the image and regroups every row.  (faiss.Kmeans.train has no such state: every train() call
 * of the look-alike goes through at_kmeans_begin.) */
int at_kmeans_invalidate(at_kmeans *km);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU exchange step of a Lloyd iteration over peer memory (one process per GPU, one node).
 * The reference has no counterpart (faiss.Kmeans(gpu=True) shards inside one process); in this library the per-iteration
 * exchange is the sum over ranks of the exact int64 accumulator.  An at_peer owns a window of
 * device memory that the other ranks map through CUDA IPC; at_peer_reduce is ONE kernel that signals, waits for every
 * rank's window and adds them up over NVLink in rank order (integers: the total is identical on every rank and for every
 * rank count).  Protocol per iteration:
 *     at_kmeans_accumulate(km, x, n, 0, algo, at_peer_local_buffer(peer), labels, stream);
 *     at_peer_reduce(peer, stream);
 *     at_kmeans_finalize(km, at_peer_total(peer), n_total, stats, stream);
 * Set-up (host, once): every rank creates its at_peer, exports a 64-byte handle, the caller exchanges the handles (any
 * transport: torch.distributed all_gather in the Python layer), every rank imports the others' and calls at_peer_connect
 * after a barrier.  A missing rank makes the kernel give up after ~4 s of SM clock; at_peer_status then reports it.
 * ---------------------------------------------------------------------------------------------- */
#define AT_PEER_MAX 16
#define AT_PEER_HANDLE_BYTES 64
typedef struct at_peer at_peer;
int at_peer_create(int rank, int world, int64_t words, at_peer **peer);
int at_peer_export(const at_peer *peer, void *handle64);
int at_peer_import(at_peer *peer, int peer_rank, const void *handle64);
int at_peer_connect(at_peer *peer);
int at_peer_destroy(at_peer *peer);
/* Device pointer of the window buffer the NEXT at_peer_reduce will read (words int64). */
int64_t *at_peer_local_buffer(at_peer *peer);
/* Device pointer of the all-rank total written by the last at_peer_reduce (words int64). */
const int64_t *at_peer_total(const at_peer *peer);
int at_peer_reduce(at_peer *peer, void *stream);
/* Synchronises the stream and returns AT_ERR_CUDA if a wait timed out since creation. */
int at_peer_status(at_peer *peer, void *stream);

/* faiss::rand_perm(perm, n, seed) (faiss/utils/random.cpp): forward Fisher-Yates driven by std::mt19937(seed),
 * i2 = i + mt() % (n - i).  HOST function, HOST pointer.  Used for FAISS's training-set subsample (seed 1234)
 * and random-point initialisation (seed 1235). */
int at_rand_perm_host(int32_t *perm, int64_t n, int64_t seed);

/* ------------------------------------------------------------------------------------------------
 * 16-bit PCM -> fp32 waveform (the step before the path: processors/spectrogram_generator.py:97-107,
 * preprocess_waveform -> torchaudio.load, which yields sample / 32768 for a 16-bit file).  Lets a caller ship
 * the decoder's native int16 samples over PCIe (half the bytes) and widen them on the device; device pointers.
 * ---------------------------------------------------------------------------------------------- */
int at_pcm16_to_f32(const int16_t *pcm, int64_t n, float *out, void *stream);

/* Channel mean + sample-rate conversion of one decoded clip: SpectrogramGenerator.convert_to_mono + .resample
 * (processors/spectrogram_generator.py:109-121) = torch.mean(dim=0) then torchaudio.transforms.Resample(orig, new) with
 * its defaults (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99).  A plan holds the filter bank of one
 * (orig_freq, new_freq) pair (the reference rebuilds it for every clip).  wave: device fp32 [channels][n_in]
 * (torchaudio.load's layout); out: device fp32 [at_resample_out_len(plan, n_in)] = ceil(new * n_in / orig) samples. */
typedef struct at_resample_plan at_resample_plan;
int at_resample_plan_create(int orig_freq, int new_freq, at_resample_plan **plan);
/* HOST function, HOST pointers, no device needed: the filter bank a plan for this rate pair uses, [*phases][*taps]
 * floats (torchaudio's _get_sinc_resample_kernel); bank may be NULL to query the sizes. */
int at_resample_bank_host(int orig_freq, int new_freq, float *bank, int *phases, int *taps, int *width);
int at_resample_plan_destroy(at_resample_plan *plan);
int64_t at_resample_out_len(const at_resample_plan *plan, int64_t n_in);
int at_resample_mono(at_resample_plan *plan, const float *wave, int channels, int64_t n_in, float *out, void *stream);
/* B clips of the same length in one launch: wave [B][channels][n_in] -> out [B][out_len]. */
int at_resample_mono_batch(at_resample_plan *plan, const float *wave, int channels, int64_t n_in, int B, float *out,
                           void *stream);

/* ------------------------------------------------------------------------------------------------
 * Token histogram (SpecTokenizer.analyze_tokens' Counter) and int32 -> int64 widening
 * ---------------------------------------------------------------------------------------------- */
int at_bincount(const int32_t *labels, int64_t n, int k, int64_t *counts, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Token consumer (the step after the path): batch assembly of TokenizedSpecDataset.__getitem__ + collate_fn
 * (datasets/tokenized_spec_dataset.py:52-76, datasets/data_loader_creator.py:17-34) from a device-resident token store.
 * tokens: flat int32 / int64 array (token_bytes 4 / 8) of all clips, clip c at [offsets[c], offsets[c+1]); idx[B]: the clips
 * of the batch.  out (B, t_max) int64 = pad_sequence(..., batch_first=True, padding_value=0); mask (B, t_max) float (may
 * be NULL): 1 on tokens / 0 on padding, or all ones when mask_all_ones != 0 -- which is what the reference's collate_fn
 * produces, because it builds the masks from the already padded matrix (:70-74).  Device pointers.
 * ---------------------------------------------------------------------------------------------- */
int at_tokens_collate(const void *tokens, int token_bytes, const int64_t *offsets, const int64_t *idx, int B, int t_max,
                      int mask_all_ones, int64_t *out, float *mask, void *stream);
/* labels (B, num_classes) float: multi-hot of label_ids[label_offsets[c] .. label_offsets[c+1]) for each clip of the batch
 * (__getitem__'s labels[label_indices] = 1.0 + collate_fn's torch.stack). */
int at_tokens_multihot(const int32_t *label_ids, const int64_t *label_offsets, const int64_t *idx, int B, int num_classes,
                       float *labels, void *stream);
/* Longest sequence of the batch -> device int32 (sizes the padded matrix). */
int at_tokens_batch_max_len(const int64_t *offsets, const int64_t *idx, int B, int32_t *out_dev, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Synthetic clips (bit-identical to oracle/synth_ref.py given the same sine table)
 * ---------------------------------------------------------------------------------------------- */
int at_synth_clips(uint32_t seed, int64_t first_index, int64_t count, int64_t n_samples,
                   const int16_t *sine_table4096, float *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* AUDIO_TOKENS_B200_H */
