/* k4b_hamm.h - C ABI of the B200-native K-mer Hamming-distance engine.
 *
 * This is the drop-in boundary for the `ngskit4b hammings` hot path.  The reference has no
 * FFI seam of its own; the seam sits exactly where the reference hands its concatenated
 * genome and result array to the worker pool:
 *   - exhaustive (-m1/-m2): ngskit4b/hammings.cpp:2740-2867 (thread params, ThreadedGHamDist,
 *     GHamDistWatson :3183-3287, GHamDistCrick :3300-3489, per-thread min-merge :2855-2867)
 *   - targeted  (-m0 -I):   ngskit4b/hammings.cpp:1691-1694 (RestrictedHammingThread calling
 *     CSfxArray::LocateHammings, libkit4b/SfxArray.cpp:4227-4331)
 * Plain pointers and sizes only; caller owns every host buffer; the library copies what it
 * needs to the device.  All entry points return 0 (eBSFSuccess, libkit4b/ErrorCodes.h:16) on
 * success or a negative teBSFrsltCodes-style value; they never throw and never exit().
 * There is no CPU fallback: without a CUDA device every compute entry point fails with
 * K4B_ERR_NODEVICE.
 */
#ifndef K4B_HAMM_H
#define K4B_HAMM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* result codes: values follow libkit4b/ErrorCodes.h:15-30 where a counterpart exists */
#define K4B_OK 0
#define K4B_ERR_PARAMS (-100)   /* eBSFerrParams */
#define K4B_ERR_MEM (-95)       /* eBSFerrMem */
#define K4B_ERR_NODEVICE (-1000) /* no usable CUDA device / library built without a device */
#define K4B_ERR_CUDA (-1001)    /* a CUDA runtime call failed; see k4b_last_error() */
#define K4B_ERR_NCCL (-1002)    /* a NCCL call failed; see k4b_last_error() */
#define K4B_ERR_UNSUPPORTED (-1003)

/* base codes of the concatenated arrays: libkit4b/commdefs.h:76-91 */
#define K4B_BASE_EOS 7u
#define K4B_BASE_EOG 0x0fu
#define K4B_RPT_MASK 0x08u

/* Limits mirrored from the reference CLI: ngskit4b/hammings.cpp:36-42 */
#define K4B_MIN_K 10u
#define K4B_MAX_K 5000u

/* ---- lifetime -------------------------------------------------------------------------- */

/* Creates contexts and streams on n_gpus devices (0 = all visible) and, when more than one
 * device is used, a single-process NCCL communicator (ncclCommInitAll) for the broadcast of the
 * packed sets and the all_reduce(min) of the per-device minima.  Replaces the reference's pthread
 * pool set-up (hammings.cpp:2752-2780). */
int k4b_gpu_init(int n_gpus, const int *device_ids /* nullable */);
void k4b_gpu_shutdown(void);
/* number of devices in use after k4b_gpu_init (0 before) */
int k4b_gpu_count(void);
/* thread-local, never NULL */
const char *k4b_last_error(void);

/* ---- exhaustive all-vs-all (-m1) --------------------------------------------------------- */

/* Replaces ThreadedGHamDist + GHamDistWatson/Crick + merge (hammings.cpp:883-939, 3183-3489,
 * 2855-2867).
 *   concat       == &m_pGenomeSeq[1] (hammings.cpp:924): codes 0..6, eBaseEOS (7) between
 *                   chromosomes, no leading/trailing EOG; soft-mask bit already cleared
 *   concat_len   == m_GenomeLen-2
 *   K            10..5000
 *   both_strands -c
 *   sweep_start/sweep_end  SSeqStart/SSeqEnd after clamping (hammings.cpp:2690-2706);
 *                   defaults 1 and m_GenomeLen (= concat_len+2) select every pair
 *   out_min      concat_len entries PRE-FILLED by the caller (K+1, hammings.cpp:3120-3122);
 *                   lowered (min) at valid K-mer start positions only
 * Engine and multi-GPU partition (all k4b_gpu_init devices are used, one NCCL broadcast of the
 * packed set first): a full sweep of >= 200 kb runs on the diagonal-band engine - the PAIR MATRIX
 * is partitioned over the devices, every device keeps a complete minima array and the arrays meet
 * in ncclAllReduce(min) after every slab; sweep sub-ranges and small inputs run on the POPC
 * all-pairs engine with the QUERIES sharded and the per-device minima concatenated on the host. */
int k4b_hamm_exhaustive(const uint8_t *concat, uint32_t concat_len, uint32_t K, int both_strands,
                        uint32_t sweep_start, uint32_t sweep_end, uint16_t *out_min);

/* Same computation restricted to the query K-mers whose flat start position lies in
 * [q_begin, q_end): the multi-GPU / multi-process shard unit (one process per GPU under
 * torchrun calls this with its own range; targets are always the whole concat). Runs on the
 * first initialised device of this process. */
int k4b_hamm_exhaustive_shard(const uint8_t *concat, uint32_t concat_len, uint32_t K,
                              int both_strands, uint32_t q_begin, uint32_t q_end,
                              uint16_t *out_min);

/* ---- targeted probes-vs-assembly (-m0 -I) -------------------------------------------------- */

/* Replaces RestrictedHammingThread + CSfxArray::LocateHammings (hammings.cpp:1593-1708,
 * SfxArray.cpp:4227-4627).  Whole probe sets of pure ACGT (cores of >= 6 bases) run on the
 * seed-and-verify engine - the reference's pigeonhole search on a bucket index, without its depth
 * cut-offs, each device building and joining its share of the index buckets; wildcard probe
 * K-mers, probe sub-ranges and short cores run on the exact brute-force engines (bands / POPC).
 *   target_concat/target_len  the .sfx sequence area: each entry's bases followed by EOS
 *                             (SfxArray.cpp:1746-1750)
 *   probe_concat/probe_len    probes in the LoadGenome layout (hammings.cpp:2201-2310); NULL
 *                             = the probes are the K-mers of the target itself
 *                             (CSfxArray::LocateSfxHammings, SfxArray.cpp:4107-4220: exact
 *                             sense-strand self hits are skipped, results capped at 20);
 *                             out_h then has target_len entries
 *   K 10..500, R 1..10; result per probe K-mer start = min(true both-strand minimum,
 *   K/(K/(R+1))) i.e. the reference's "not found" value (SfxArray.cpp:4462-4463); probe
 *   symbols >= N are wildcards against target ACGT, > 4 of them report 0 (:4266-4326)
 *   out_h  probe_len entries PRE-FILLED 0xFF; written at valid probe K-mer starts */
int k4b_hamm_targeted(const uint8_t *target_concat, uint64_t target_len,
                      const uint8_t *probe_concat, uint32_t probe_len, uint32_t K, int R,
                      int both_strands, uint32_t q_begin, uint32_t q_end, uint8_t *out_h);
/* The NULL-probe form (K-mers of the indexed assembly against itself) with the reference's -z
 * IntraInterBoth option (hammings.cpp:228; SfxArray.cpp:4421-4426, :4597-4601): 0 = both,
 * 1 = intra only, 2 = inter only.  As in the reference it filters ONLY exact (0-mismatch)
 * sense-strand hits: such a hit counts only if it lies in the same (1) / in another (2) entry
 * than the probe K-mer; entries are the EOS-terminated runs of target_concat. */
int k4b_hamm_targeted_z(const uint8_t *target_concat, uint64_t target_len, uint32_t K, int R,
                        int both_strands, int intra_inter_both, uint32_t q_begin, uint32_t q_end,
                        uint8_t *out_h);

/* ---- device-resident API (inputs already in HBM; used by bench.py and by ranks that receive
 *      the packed target set over NCCL instead of packing it themselves) --------------------- */

typedef struct k4b_packed k4b_packed; /* opaque: bit-plane packed sequence set on one device */

/* bytes of one device-resident packed image for a concat of this length (for NCCL broadcast) */
size_t k4b_packed_image_bytes(uint32_t concat_len);
/* Packs a HOST concat (H2D copy + pack kernel on the current device of this process).  An image is
 * 6 arrays: bit-planes 0-2, the valid-K-mer-start plane, and a 2-bit-per-base code array (2 arrays). */
int k4b_pack_host(const uint8_t *concat, uint32_t concat_len, uint32_t K, k4b_packed **out);
/* Packs a DEVICE-resident concat (16-byte aligned; pack + valid-start kernels only). */
int k4b_pack_device(const void *d_concat, uint32_t concat_len, uint32_t K, void *stream,
                    k4b_packed **out);
/* Same, writing the image into a caller-owned device buffer of k4b_packed_image_bytes() bytes
 * (e.g. a buffer that is then broadcast to the other ranks over NCCL). */
int k4b_pack_device_into(const void *d_concat, uint32_t concat_len, uint32_t K, void *d_image,
                         size_t image_bytes, void *stream, k4b_packed **out);
/* device pointer + size of the packed image (planes+valid bits), e.g. for ncclBroadcast */
void *k4b_packed_image_ptr(k4b_packed *p);
size_t k4b_packed_image_size(k4b_packed *p);
/* Rebuilds a handle around an image received from another rank (image stays owned by caller).*/
int k4b_packed_from_image(void *d_image, size_t image_bytes, uint32_t concat_len, uint32_t K,
                          int has_non_acgt, k4b_packed **out);
int k4b_packed_has_non_acgt(k4b_packed *p);
uint64_t k4b_packed_num_kmers(k4b_packed *p);
void k4b_packed_free(k4b_packed *p);

/* All-pairs minimum on device-resident data: queries = K-mers of `queries` starting in
 * [q_begin,q_end), targets = all K-mers of `targets`.  self_exclude!=0 skips the pair
 * (query position == target position) on the forward strand (Watson offsets start at 1,
 * hammings.cpp:2692; the reverse-complement self pair is kept, :3300-3489).
 * d_out_min: DEVICE uint16[q_end-q_begin]; positions that are not valid K-mer starts get
 * K+1.  clamp>0 selects the targeted (-m0) rules: results are capped at clamp, query symbols
 * >= N act as wildcards and a K-mer with more than 4 of them reports 0
 * (SfxArray.cpp:4266-4326).  Asynchronous on `stream`.
 * *launches (nullable) receives the number of kernels enqueued. */
int k4b_allpairs_min_device(k4b_packed *queries, k4b_packed *targets, int both_strands,
                            int self_exclude, uint32_t q_begin, uint32_t q_end, uint32_t clamp,
                            uint16_t *d_out_min, void *stream, int *launches);
/* duration in ms of the most recent allpairs kernel enqueued by this thread, measured with
 * CUDA events on the launching stream (blocks until that kernel has finished) */
float k4b_last_kernel_ms(void);

/* ---- engine selection and the diagonal-band engine on device-resident data ------------------- */
/* Two exact engines produce the exhaustive minima: 1 = POPC all-pairs (queries in registers,
 * XOR / fold / POPC per 32 bases), 2 = diagonal bands (the reference's O(1)-per-pair sliding
 * recurrence, 32 diagonals per thread in bit-sliced counters, every cell serving both K-mers
 * of the pair).  0 = automatic: exhaustive - bands for full sweeps of >= 200 kb, all-pairs for
 * smaller inputs, sweep sub-ranges (-b/-B, -m2) and query shards; targeted - the seed-and-verify
 * engine (3) for pure-ACGT probe sets with cores of >= 6 bases, else bands, all-pairs for probe
 * sub-ranges and wildcard probe K-mers.  Env: K4B_ENGINE=popc|diag|seed. */
int k4b_set_engine(int engine);
int k4b_get_engine(void);
/* d_best: DEVICE uint32[len] running minima.  init fills K+1; k4b_exhaustive_diag_device lowers
 * them for part `part` of `nparts` of the pair matrix (parts are independent and combine by an
 * element-wise minimum, e.g. ncclAllReduce(min)); finalize converts to the uint16 layout of
 * k4b_allpairs_min_device (K+1 where no K-mer starts). All asynchronous on `stream`. */
int k4b_best_init_device(uint32_t *d_best, uint32_t n, uint32_t K, void *stream);
int k4b_exhaustive_diag_device(k4b_packed *g, int both_strands, uint32_t part, uint32_t nparts,
                               uint32_t *d_best, void *stream, int *launches);
int k4b_best_finalize_device(k4b_packed *g, const uint32_t *d_best, uint16_t *d_out_min, void *stream);
/* The two phases of k4b_exhaustive_diag_device, for multi-GPU runs that shard both:
 * bootstrap = K-mers starting in [q_begin,q_end) against a small target sample (POPC engine);
 * bands = part `part` of `nparts` of the pair matrix.  Combine d_best across ranks by an
 * element-wise minimum after each phase. */
int k4b_diag_bootstrap_device(k4b_packed *g, int both_strands, uint32_t q_begin, uint32_t q_end,
                              uint32_t *d_best, void *stream);
int k4b_diag_bands_device(k4b_packed *g, int both_strands, uint32_t part, uint32_t nparts,
                          uint32_t *d_best, void *stream, int *launches);
/* The bands run in SLABS (thresholds = block maxima of d_best are refreshed before each; the
 * first slabs are small so that thresholds tighten early).  k4b_diag_slab_count gives the number
 * of slabs for this image and nparts; k4b_diag_slabs_device runs slabs [slab_begin, slab_end) of
 * one part, so that a multi-GPU driver can combine d_best across ranks (element-wise minimum)
 * BETWEEN slabs and every rank thresholds against what all ranks have found so far.
 * k4b_diag_bands_device == all slabs in one call. */
int k4b_diag_slab_count(k4b_packed *g, int both_strands, uint32_t nparts, uint32_t *n_slabs);
int k4b_diag_slabs_device(k4b_packed *g, int both_strands, uint32_t part, uint32_t nparts,
                          uint32_t slab_begin, uint32_t slab_end, uint32_t *d_best, void *stream,
                          int *launches);

/* Counter widths of the most recent band run (k4b_diag_bands_device / _slabs_device) of this thread: planes of the full
 * and of the narrow kernel instance (0 = none), number of slabs, and how many slabs ran narrow
 * (chosen on the device from the slab's largest threshold).  Blocks until that run finished. */
int k4b_last_diag_info(uint32_t *np_full, uint32_t *np_small, uint32_t *slabs, uint32_t *narrow_slabs);

/* Targeted (probes vs assembly, -m0 -I) on the band engine: rows = probe K-mers, columns = target
 * K-mers, fixed threshold = clamp (the "not found" value), targeted wildcard rules.  d_best:
 * DEVICE uint32[probe len] initialised by k4b_best_init_device; part/nparts as above. */
int k4b_targeted_diag_device(k4b_packed *probes, k4b_packed *targets, int both_strands, uint32_t clamp,
                             uint32_t part, uint32_t nparts, uint32_t *d_best, void *stream, int *launches);
/* Targeted by seed-and-verify (the reference's pigeonhole search, SfxArray.cpp:4462-4581, without
 * its depth cut-offs): every core of core_len bases of a probe K-mer is looked up exactly in a
 * bucket index of the target's cores and each occurrence is verified over the full K bases.
 * Exact for distances < clamp, which needs clamp <= K/core_len.  Probe K-mers that hold N / InDel
 * (wildcards) are skipped - k4b_hamm_targeted answers those with k4b_allpairs_min_device.  Probe
 * K-mers starting in [q_begin, q_end); d_best as above.
 * Device memory (stream-ordered pool, released when the call's work is done): 16 bytes per target
 * base for the index, and as much again while it is built when core_len <= 8 (two partition passes
 * through a scratch array; without room for it, or with K4B_SEED_INDEX=0, the entries are placed
 * one by one instead: slower, same index). */
int k4b_targeted_seed_device(k4b_packed *probes, k4b_packed *targets, int both_strands, uint32_t clamp,
                             uint32_t core_len, uint32_t q_begin, uint32_t q_end, uint32_t *d_best,
                             void *stream, int *launches);
/* Same engine, split by INDEX BUCKETS instead of probes: part `part` of `nparts` indexes only its
 * share of the core buckets (1/nparts of the index build) and answers every probe K-mer for those
 * buckets.  Parts combine by an element-wise minimum of d_best (e.g. ncclAllReduce(min)). */
int k4b_targeted_seed_part_device(k4b_packed *probes, k4b_packed *targets, int both_strands, uint32_t clamp,
                                  uint32_t core_len, uint32_t part, uint32_t nparts, uint32_t *d_best,
                                  void *stream, int *launches);
/* Work of the most recent seed-engine call of this thread: bucket entries tested by the query
 * kernel (16-byte entries) and cores held by the index.  Blocks until that call finished. */
int k4b_last_seed_info(uint64_t *occurrences, uint64_t *indexed_cores);
int k4b_targeted_finalize_device(k4b_packed *probes, const uint32_t *d_best, uint32_t clamp,
                                 uint16_t *d_out_min, void *stream);

/* ---- where the reference's own answer may differ ------------------------------------------------
 * The reference's pigeonhole search walks at most MaxCoreDepth suffix-array entries per core and gives
 * a core up when it has more copies than that (SfxArray.cpp:4480-4494; MaxCoreDepth by -s,
 * hammings.cpp:2366-2386, times a per-level multiplier, SfxArray.cpp:4297-4311); on repeat-rich
 * assemblies it then reports a LARGER value than the true minimum.  The seed engine has no such cut.
 * k4b_set_reference_sensitivity tells it the caller's -s (0 default, 1 more, 2 ultra, 3 less);
 * k4b_last_depth_cut returns, for the most recent seed-engine k4b_hamm_targeted[_z] call of this thread,
 * how many probe K-mers (a) hold a core occurring more than *max_copies times in the assembly (the cap of
 * the last cascade level) and (b) were answered below the "not found" value - the only places where a
 * file diff against the reference can show a difference.  Device API: k4b_seed_watch_depth arms the next
 * k4b_targeted_seed[_part]_device call of this thread (d_deep: one byte per probe position, zeroed by the
 * caller, set to 1 for flagged positions; pass 0 / NULL to disarm). */
int k4b_set_reference_sensitivity(int sensitivity);
int k4b_last_depth_cut(uint64_t *probe_kmers, uint32_t *max_copies);
int k4b_seed_watch_depth(uint32_t max_copies, uint8_t *d_deep);

/* ---- distribution of the minima -------------------------------------------------------------
 * The table `hammings` logs after an exhaustive run (hammings.cpp:2939-2962) and the downstream
 * HammingDist tool tabulates from the CSV (HammingDist/HammingDist.cpp:371-705): hist[d] = number of
 * positions whose minimum is d (d = 0..K); hist[K+1] collects the positions where no K-mer starts
 * (they hold K+1).  hist has K+2 entries.  Host-buffer form and device form (d_hist is zeroed by the
 * callee; asynchronous on `stream`). */
int k4b_hamm_histogram(const uint16_t *min, uint32_t n, uint32_t K, uint64_t *hist);
int k4b_histogram_device(const uint16_t *d_min, uint32_t n, uint32_t K, unsigned long long *d_hist, void *stream);

/* ---- integer-pipe roofline microbenchmark ----------------------------------- */
/* which: 0 POPC only, 1 LOP3 only, 2 engine mix (2 LOP3 + 1 POPC + min), 3 IADD3 only.
 * Returns giga warp-lane-ops per second (ops/s / 1e9) on the current device in *gops. */
int k4b_microbench_intpipe(int which, int iters, double *gops);

#ifdef __cplusplus
}
#endif
#endif /* K4B_HAMM_H */
