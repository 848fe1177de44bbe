/*
 * b200aln.h -- C ABI of the B200 batched overlap aligner (the step BEFORE the POA path).
 *
 * What racon binds for this step is the cudaaligner `Aligner` / `Alignment` interface, used by its adapter
 * racon::CUDABatchAligner (/root/reference/src/cuda/cudaaligner.{hpp,cpp}); every entry point below names the member
 * it replaces.  Reference paths are relative to /root/reference; "aligner.hpp" / "alignment.hpp" / "cudaaligner.hpp"
 * are vendor/GenomeWorks/cudaaligner/include/claraparabricks/genomeworks/cudaaligner/{aligner,alignment,cudaaligner}.hpp.
 *
 * SEMANTICS = racon's CPU path: the alignment edlib returns for
 *     edlibAlign(q, t, {k = -1, EDLIB_MODE_NW, EDLIB_TASK_PATH}) + edlibAlignmentToCigar(EDLIB_CIGAR_STANDARD)
 * (src/overlap.cpp:205-224) -- same edit distance AND the same choice among the optimal alignments, so that racon's
 * breaking points, windows and consensus stay those of `racon -t` (cudaaligner returns *an* optimal or, past its band,
 * a sub-optimal alignment).  No band that can cost optimality (the engine confines itself to bands only where that is
 * provably or verifiably exact), no length limit other than memory: nothing is left to a CPU aligner.
 *
 * Plain pointers and sizes only; every function returns a b200aln_status (= cudaaligner::StatusType values,
 * cudaaligner.hpp:34-42) unless stated otherwise.  A batch belongs to one device and one stream and is not thread-safe;
 * different batches may be driven from different threads (racon: one thread per batch, cudapolisher.cpp:139-174).
 */
#ifndef B200ALN_H
#define B200ALN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum b200aln_status { /* cudaaligner.hpp:34-42 */
    B200ALN_SUCCESS = 0,
    B200ALN_UNINITIALIZED = 1,
    B200ALN_EXCEEDED_MAX_ALIGNMENTS = 2,           /* the batch's memory budget is used up: align, reset, add again */
    B200ALN_EXCEEDED_MAX_LENGTH = 3,               /* this one pair can never fit the budget, even alone            */
    B200ALN_EXCEEDED_MAX_ALIGNMENT_DIFFERENCE = 4, /* never produced (no lossy band); kept for the numbering        */
    B200ALN_GENERIC_ERROR = 5,
    B200ALN_INVALID_ARGUMENT = 6,
    B200ALN_CUDA_ERROR = 7
};

/* edit operations of an alignment, per position: edlib's EDLIB_EDOP_* codes (vendor/edlib/edlib/include/edlib.h) */
enum b200aln_op {
    B200ALN_OP_MATCH = 0,
    B200ALN_OP_INSERT = 1,  /* a query character alone  ('I') */
    B200ALN_OP_DELETE = 2,  /* a target character alone ('D') */
    B200ALN_OP_MISMATCH = 3
};

typedef struct b200aln_batch b200aln_batch;

/* cudaaligner::Init() (cudaaligner.hpp:61): 0 when a CUDA device is usable. */
int32_t b200aln_init(void);

/* create_aligner(AlignmentType::global_alignment, max_bandwidth, stream, device_id, max_device_memory)
 * (aligner.hpp:121-132; racon: src/cuda/cudaaligner.cpp:24-45).  `stream` is a cudaStream_t or NULL (the batch then
 * owns one).  max_gpu_mem <= 0: 90 % of the device's free memory.  max_bandwidth is accepted for signature parity and
 * ignored: the engine never trades optimality for a band (see b200aln_batch_set_band_guess). */
int32_t b200aln_batch_create(int32_t device_id, void* stream, int64_t max_gpu_mem, int32_t max_bandwidth,
                             b200aln_batch** out);

/* How a batch bands the TOP level of its alignments.  Sub-problems below the top know their optimum exactly and run in a
 * band of that width; the top one's is unknown, so the host guesses `permille` edits per 1000 characters of the longer
 * sequence and the kernels verify (an alignment whose distance exceeds the guess is redone without a band: results never
 * depend on the guess, only speed does).  -1 (default): learnt from the batch's previous align_all (1.25 x the rate 90 % of
 * its alignments stayed under; a fresh batch does not guess); 0: never; > 0: that rate.  max_bandwidth of create_aligner
 * is NOT this: cudaaligner's band silently yields sub-optimal alignments, this one cannot. */
int32_t b200aln_batch_set_band_guess(b200aln_batch* b, int32_t permille);

/* Aligner::add_alignment(query, query_length, target, target_length) (aligner.hpp:73-74).  The bytes are copied.
 * NOTE the argument order is edlib's (query = the read segment, target = the contig segment, src/overlap.cpp:205-209);
 * racon's adapter swaps them for cudaaligner (cudaaligner.cpp:60-63) -- the C++ shim keeps that call site intact by
 * swapping back.  Returns B200ALN_EXCEEDED_MAX_ALIGNMENTS when the batch is full (nothing was added). */
int32_t b200aln_batch_add_alignment(b200aln_batch* b, const char* query, int32_t query_length, const char* target,
                                    int32_t target_length);

/* ---- breaking points (extension; what racon does with the CIGAR next: Overlap::find_breaking_points_from_cigar,
 * src/overlap.cpp:226-290, called per overlap on CPU threads after generate_cigar_strings, cudapolisher.cpp:190-214).
 * window_length > 0 (racon -w, default 500) makes align_all also form every overlap's breaking points on the device --
 * per window of the contig that the overlap touches and that holds an aligned column: its first one as (t, q), its last
 * one as (t + 1, q + 1), contig / read coordinates -- and download them (16 bytes a window).  skip_cigars != 0: the CIGAR
 * text is then neither formed nor downloaded.  Call on an empty batch (after create or reset). */
int32_t b200aln_batch_set_window_length(b200aln_batch* b, int32_t window_length, int32_t skip_cigars);

/* add_alignment with the segments' coordinates: q_first = where the read segment starts in the (strand-adjusted) read,
 * i.e. strand ? q_length - q_end : q_begin (overlap.cpp:241); t_begin = where the contig segment starts (:242). */
int32_t b200aln_batch_add_overlap(b200aln_batch* b, const char* query, int32_t query_length, const char* target,
                                  int32_t target_length, int32_t q_first, int32_t t_begin);
/* columnar (q_first / t_begin nullable = 0) */
int32_t b200aln_batch_add_overlaps(b200aln_batch* b, int64_t n, const uint8_t* q_bases, const int64_t* q_off,
                                   const uint8_t* t_bases, const int64_t* t_off, const int32_t* q_first,
                                   const int32_t* t_begin, int64_t* n_added);

/* The same without the staging copy: the batch (which must be empty) REFERENCES the caller's columnar buffers for pairs
 * 0 .. *n_added - 1 and uploads those byte ranges straight from them in align_all; the buffers must stay valid and
 * unchanged until align_all has returned.  Page-lock them once (b200aln_host_register, or any pinned allocation) and the
 * upload runs at full PCIe speed. */
int32_t b200aln_batch_add_overlaps_view(b200aln_batch* b, int64_t n, const uint8_t* q_bases, const int64_t* q_off,
                                        const uint8_t* t_bases, const int64_t* t_off, const int32_t* q_first,
                                        const int32_t* t_begin, int64_t* n_added);
int32_t b200aln_host_register(const void* p, int64_t bytes);   /* cudaHostRegister (portable) */
int32_t b200aln_host_unregister(const void* p);

/* After sync: overlap k's breaking points are count[k] pairs (t, q) of uint32 at points + 2 * off[k] (always an even
 * number: first / last per window, in window order = Overlap::breaking_points(), overlap.hpp:68-70). */
int32_t b200aln_batch_get_breaking_points(const b200aln_batch* b, const uint32_t** points, const int64_t** off,
                                          const int32_t** count);

/* Columnar form of add_alignment for callers that hold their segments back to back (pair k = q_bases[q_off[k] ..
 * q_off[k + 1]) against t_bases[t_off[k] .. t_off[k + 1])): adds pairs 0, 1, .. until the batch is full; *n_added says
 * how many went in (B200ALN_SUCCESS if at least one did, or n == 0). */
int32_t b200aln_batch_add_alignments(b200aln_batch* b, int64_t n, const uint8_t* q_bases, const int64_t* q_off,
                                     const uint8_t* t_bases, const int64_t* t_off, int64_t* n_added);

/* Aligner::align_all() (aligner.hpp:56): upload, the level-synchronous Hirschberg recursion, the leaf tracebacks, run
 * and CIGAR formation on the device, and the asynchronous download of the compact results (the CIGAR bytes). */
int32_t b200aln_batch_align_all(b200aln_batch* b);

/* Aligner::sync_alignments() (aligner.hpp:62): blocks until the results are on the host. */
int32_t b200aln_batch_sync(b200aln_batch* b);

/* Aligner::get_alignments().size() (aligner.hpp:79) */
int32_t b200aln_batch_num_alignments(const b200aln_batch* b);

/* One alignment after sync (Alignment::get_status / get_edit_distance / get_alignment, alignment.hpp:87-101).
 * *runs points at n_runs words `start << 2 | op` -- run k covers operations [start_k, start_{k+1}) and the last one ends
 * at *n_ops; valid until reset/destroy.  Any out pointer may be NULL; asking for `runs` fetches the batch's run starts
 * from the device on first use (racon only needs the CIGARs, which align_all already brought back). */
int32_t b200aln_batch_get_alignment(const b200aln_batch* b, int32_t index, const uint32_t** runs, int32_t* n_runs,
                                    int32_t* n_ops, int32_t* edit_distance, int32_t* status);

/* All CIGARs of the batch after sync, in place (what CUDABatchAligner::generate_cigar_strings walks,
 * cudaaligner.cpp:88-103): alignment k's string is text[off[k] .. off[k] + len[k]) followed by a 0; edit_distance[k],
 * status[k] per alignment.  Valid until reset/destroy; any out pointer may be NULL. */
int32_t b200aln_batch_get_cigars(const b200aln_batch* b, const char** text, const int64_t** off, const int32_t** len,
                                 const int32_t** edit_distance, const int32_t** status);

/* Alignment::convert_to_cigar() (alignment.hpp:72) in edlib's EDLIB_CIGAR_STANDARD spelling ("12M1I3M2D...").
 * Returns the string's length (without the terminating 0); writes at most cap bytes including the 0; out may be NULL
 * to ask for the length; negative = -status. */
int64_t b200aln_batch_get_cigar(const b200aln_batch* b, int32_t index, char* out, int64_t cap);

/* Alignment::get_alignment() (alignment.hpp:93) expanded to one b200aln_op per position.  Returns the number of
 * operations, writes at most cap of them; negative = -status. */
int64_t b200aln_batch_get_ops(const b200aln_batch* b, int32_t index, uint8_t* out, int64_t cap);

/* Aligner::reset() (aligner.hpp:82): forget the alignments, keep the device memory. */
int32_t b200aln_batch_reset(b200aln_batch* b);
void b200aln_batch_destroy(b200aln_batch* b);

typedef struct b200aln_batch_info {
    int32_t device_id;
    int32_t n_slots;          /* resident warps = per-warp workspaces                         */
    int32_t levels;           /* Hirschberg levels of the last align_all                      */
    int32_t kernel_launches;  /* launches of the last align_all                               */
    int32_t team_launches;    /* of them: team launches (a block of warps per tall sub-problem) */
    int32_t n_team_blocks;    /* resident team blocks = team workspaces                        */
    int64_t n_open;           /* sub-problems split in the last align_all                     */
    int64_t n_leaves;         /* sub-problems traced back directly                            */
    int64_t cells;            /* distance-matrix cells actually computed (splits + leaves, inside their bands) */
    int64_t h2d_bytes, d2h_bytes;
    float kernel_ms;          /* device time of the last align_all's launches (CUDA events)   */
} b200aln_batch_info;
int32_t b200aln_batch_get_info(const b200aln_batch* b, b200aln_batch_info* info);

const char* b200aln_status_string(int32_t status);

/* Bulk form used by the bench and by callers holding columnar data: aligns n pairs (sequences back to back in
 * q_bases / t_bases, per-pair offsets, as many batches as the budget needs) and returns per pair the edit distance and the CIGAR
 * (cigars: caller's buffer of cigar_cap bytes, strings back to back, cigar_off[n + 1]).  Returns a b200aln_status;
 * B200ALN_EXCEEDED_MAX_LENGTH when cigar_cap is too small (cigar_off[n] then holds the bytes needed). */
int32_t b200aln_align_pairs(int32_t device_id, int64_t max_gpu_mem, int64_t n, const uint8_t* q_bases,
                            const int64_t* q_off, const uint8_t* t_bases, const int64_t* t_off, int32_t* edit_distance,
                            char* cigars, int64_t cigar_cap, int64_t* cigar_off, b200aln_batch_info* info);

/* ---- aligner pool: the GPU section of CUDAPolisher::find_overlap_breaking_points (src/cuda/cudapolisher.cpp:74-214:
 * `cudaaligner_batches` batches per device, every device, one host thread per batch in a fill -> alignAll ->
 * generate_cigar_strings -> reset loop, :139-174) over columnar segments.  max_gpu_mem_per_batch <= 0: 90 % of each
 * device's free memory split between its batches (:118-123). */
typedef struct b200aln_aligner b200aln_aligner;
int32_t b200aln_aligner_create(int32_t n_devices, const int32_t* device_ids, int32_t batches_per_device,
                               int64_t max_gpu_mem_per_batch, b200aln_aligner** out);
int32_t b200aln_aligner_num_batches(const b200aln_aligner* h);
/* Aligns n pairs (pair k = q_bases[q_off[k] .. q_off[k + 1]) against t_bases[t_off[k] .. t_off[k + 1])).  Pair k's CIGAR
 * is the cigar_len[k] bytes at cigars + cigar_off[k], followed by a 0 (the strings of one batch lie together, batches in
 * completion order); edit_distance[k] its distance (nullable).  *cigar_bytes = bytes of `cigars` used or needed;
 * B200ALN_EXCEEDED_MAX_LENGTH when cigar_cap was too small (offsets of the strings that did not fit are -1).
 * info (nullable): sums over the batches (kernel_ms adds device times of batches that overlap). */
int32_t b200aln_aligner_align(b200aln_aligner* h, int64_t n, const uint8_t* q_bases, const int64_t* q_off,
                              const uint8_t* t_bases, const int64_t* t_off, int32_t* edit_distance, char* cigars,
                              int64_t cigar_cap, int64_t* cigar_off, int32_t* cigar_len, int64_t* cigar_bytes,
                              b200aln_batch_info* info);
void b200aln_aligner_destroy(b200aln_aligner* h);

#ifdef __cplusplus
}
#endif
#endif /* B200ALN_H */
