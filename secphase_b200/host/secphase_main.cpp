// secphase_main.cpp -- the `secphase` command line on top of the B200 marker-mode engine.
//
// Drop-in for the reference executable built from programs/src/secphase.c (paths relative to
// /root/reference/): same option table and defaults (secphase.c:387-449), same order-sensitive
// --hifi/--ont presets (477-504), same output files (619,650,682,715-732) and stderr lines
// (261,324-330,706-709,64-70).  What differs is the inside of the scatter loop: instead of one
// runOneThread job per read group on a pthread pool (secphase.c:303), eligible groups are packed
// into batches (libsecphase_host: BGZF/BAM decode on -@ threads) and scored on the GPU(s)
// through the C ABI of include/secphase_b200.h, three batches in flight per device.  Records of
// out.log appear in input order (the reference's order with -@1; with more workers its order is
// completion order).
//
// -w/--writeBam (secphase.c:182-189, 643-657): the device then evaluates the BAQ of every base of
// every realigned window (sp_set_write_qual) and the records of all scored read groups are written
// to <outDir>/<prefix>.quality_modified.out.bam with the modified qualities -- as SAM text, which
// is what the reference's sam_open(path, "w") produces despite the file name (sph_sam.cpp).
//
// Not available in this build (out of the hot path's scope): -v/--inputVcf variant
// mode; it stops with a message instead of being silently ignored.
//
// Extra long options (not in the reference): --gpus N (default 1; read groups are dealt to the
// devices batch by batch, results merged in input order, no collective), --batchGroups N,
// --batchMB N.
#include <getopt.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <unistd.h>

#include <atomic>
#include <cfloat>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <deque>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/secphase_b200.h"
#include "../../include/secphase_host.h"

namespace {

const char *get_timestamp() {  // common.c:12-20
    static thread_local char ts[80];
    time_t t = time(nullptr);
    struct tm tmv;
    localtime_r(&t, &tmv);
    snprintf(ts, sizeof(ts), "%04d-%02d-%02d %02d:%02d:%02d", tmv.tm_year + 1900, tmv.tm_mon + 1, tmv.tm_mday,
             tmv.tm_hour, tmv.tm_min, tmv.tm_sec);
    return ts;
}

struct option long_options[] = {{"inputBam", required_argument, nullptr, 'i'},
                                {"inputFasta", required_argument, nullptr, 'f'},
                                {"inputVcf", required_argument, nullptr, 'v'},
                                {"disableMarkerMode", no_argument, nullptr, 'M'},
                                {"baq", no_argument, nullptr, 'q'},
                                {"gapOpen", required_argument, nullptr, 'd'},
                                {"gapExt", required_argument, nullptr, 'e'},
                                {"bandwidth", required_argument, nullptr, 'b'},
                                {"consensus", no_argument, nullptr, 'c'},
                                {"indelThreshold", required_argument, nullptr, 't'},
                                {"initQ", required_argument, nullptr, 's'},
                                {"minQ", required_argument, nullptr, 'm'},
                                {"primMarginScore", required_argument, nullptr, 'p'},
                                {"primMarginRandom", required_argument, nullptr, 'r'},
                                {"minScore", required_argument, nullptr, 'n'},
                                {"hifi", no_argument, nullptr, 'x'},
                                {"ont", no_argument, nullptr, 'y'},
                                {"minVariantMargin", required_argument, nullptr, 'g'},
                                {"prefix", required_argument, nullptr, 'P'},
                                {"outDir", required_argument, nullptr, 'o'},
                                {"variantBed", required_argument, nullptr, 'B'},
                                {"minGQ", required_argument, nullptr, 'G'},
                                {"threads", required_argument, nullptr, '@'},
                                {"writeBam", no_argument, nullptr, 'w'},
                                {"flankMargin", required_argument, nullptr, 'F'},
                                {"gpus", required_argument, nullptr, 1000},
                                {"batchGroups", required_argument, nullptr, 1001},
                                {"batchMB", required_argument, nullptr, 1002},
                                {"version", no_argument, nullptr, 1003},
                                {nullptr, 0, nullptr, 0}};

void usage(const char *program) {  // secphase.c:563-599, plus the three extra options
    fprintf(stderr, "\nUsage: %s  -i <INPUT_BAM> -f <FASTA> \n", program);
    fprintf(stderr, "Options:\n");
    fprintf(stderr, "         --inputBam, -i         Input BAM file\n");
    fprintf(stderr, "         --inputFasta, -f         Input FASTA file\n");
    fprintf(stderr, "         --inputVcf, -v         Input phased VCF file (not available in the B200 build)\n");
    fprintf(stderr, "         --variantBed, -B         Input BED file for subsetting phased variants\n");
    fprintf(stderr, "         --outDir, -o         Output dir for saving outputs [Default = \"secphase_out_dir\"]\n");
    fprintf(stderr, "         --prefix, -P         Prefix of the output files [Default = \"secphase\"]\n");
    fprintf(stderr,
            "         --disableMarkerMode, -M         If alignments do not overlap with variants Secphase will not switch to marker mode\n");
    fprintf(stderr,
            "         --hifi, -x         hifi preset params (only for marker mode) [-q -c -t10 -d 1e-4 -e 0.1 -b20 -m10 -s40 -p40 -r0 -n -10] (Only one of --hifi or --ont should be enabled)\n");
    fprintf(stderr,
            "         --ont, -y        ont preset params (only for marker mode) [-q -c -t20 -d 1e-3 -e 0.1 -b20 -m10 -s20 -p20 -r0 -n -10] (Only one of --hifi or --ont should be enabled) \n");
    fprintf(stderr, "         --baq, -q         Calculate BAQ [Disabled by default]\n");
    fprintf(stderr, "         --gapOpen, -d         Gap prob [Default: 1e-4, (for ONT use 1e-2)]\n");
    fprintf(stderr, "         --gapExt, -e         Gap extension [Default: 0.1]\n");
    fprintf(stderr, "         --bandwidth, -b         DP bandwidth [Default: 20]\n");
    fprintf(stderr, "         --consensus, -c         Use consensus confident blocks [Disabled by default]\n");
    fprintf(stderr,
            "         --indelThreshold, -t         Indel size threshold for confident blocks [Default: 10 (for ONT use 20)]\n");
    fprintf(stderr,
            "         --initQ, -s         Before calculating BAQ set all base qualities to this number [Default: 40 (for ONT use 20)]\n");
    fprintf(stderr,
            "         --minQ, -m         Minimum base quality (or BAQ if -q is set) to be considered as a marker  [Default: 20 (for ONT use 10)]\n");
    fprintf(stderr,
            "         --primMarginScore, -p         Minimum margin between the consistency score of primary and secondary alignment to select the secondary alignment [Default: 40]\n");
    fprintf(stderr,
            "         --primMarginRandom, -r         Maximum margin between the consistency score of primary and secondary alignment to select one randomly [Default: 0]\n");
    fprintf(stderr,
            "         --minScore, -n         Minimum marker score of the selected secondary alignment [Default: -10]\n");
    fprintf(stderr,
            "         --minVariantMargin, -g         Minimum margin for creating blocks around phased variants [Default: 50]\n");
    fprintf(stderr, "         --minGQ, -G         Minimum genotype quality of the phased variants [Default: 10]\n");
    fprintf(stderr,
            "         --writeBam, -w         Write an output bam file with the base qualities modified by BAQ\n");
    fprintf(stderr, "         --flankMargin, -F         Margin around each marker for the BAQ windows [Default: 500]\n");
    fprintf(stderr, "         --threads, -@         Number of host threads for BAM decoding [Default: 4]\n");
    fprintf(stderr, "         --gpus         Number of GPUs to shard read groups over [Default: 1]\n");
    fprintf(stderr, "         --batchGroups         Read groups per GPU batch [Default: 4096]\n");
    fprintf(stderr, "         --batchMB         Megabytes of record data per GPU batch [Default: 256]\n");
}

// ---------------------------------------------------------------- a tiny blocking queue
template <class T>
class Queue {
public:
    void push(T v) {
        {
            std::lock_guard<std::mutex> g(mu_);
            q_.push_back(std::move(v));
        }
        cv_.notify_all();
    }
    bool pop(T &out) {  // false once closed and drained
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return !q_.empty() || closed_; });
        if (q_.empty()) return false;
        out = std::move(q_.front());
        q_.pop_front();
        return true;
    }
    int try_pop(T &out) {  // 1 got one, 0 empty right now, -1 closed and drained
        std::lock_guard<std::mutex> g(mu_);
        if (q_.empty()) return closed_ ? -1 : 0;
        out = std::move(q_.front());
        q_.pop_front();
        return 1;
    }
    void close() {
        {
            std::lock_guard<std::mutex> g(mu_);
            closed_ = true;
        }
        cv_.notify_all();
    }

private:
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<T> q_;
    bool closed_ = false;
};

struct Work {
    sph_batch *hb = nullptr;
    int64_t seq = -1;
};

// everything the ordered output stage needs from one scored batch
struct Done {
    sph_batch *hb = nullptr;
    int64_t seq = -1;
    std::vector<int32_t> group;   // [G][SP_GROUP_W]
    std::vector<double> score;    // [A]
    std::vector<int32_t> extent;  // [A][4]
    std::vector<int64_t> marker_off;
    std::vector<int32_t> marker;
    std::vector<uint8_t> baq_qual;  // -w: the records' qualities after BAQ (sp_result.baq_qual)
    int64_t hmm_instances = 0, hmm_cells = 0;
    double gpu_ms = 0, hmm_ms = 0;
    int launches = 0;
};

struct Shared {
    std::mutex mu;
    std::string error;  // first fatal error
    bool failed = false;
    // called once, by the first failing thread: closes the pipeline's queues so that every thread blocked
    // on one of them (reader on free_q, an idle GPU thread on work_q) wakes up and winds down
    std::function<void()> on_fail;
    void fail(const std::string &msg) {
        {
            std::lock_guard<std::mutex> g(mu);
            if (failed) return;
            failed = true;
            error = msg;
        }
        if (on_fail) on_fail();
    }
    bool is_failed() {
        std::lock_guard<std::mutex> g(mu);
        return failed;
    }
};

// get_best_record_index (ptAlignment.c:137-177) with the process's own rand(), used when the
// groups of several devices are merged: the draws then happen here, in input order, exactly as
// in a single-worker run of the reference.
int select_best_host(int n, const int32_t *flag, const double *score, double prim_margin, double min_score,
                     double prim_margin_random) {
    if (n == 1) return 0;
    double max_score = -DBL_MAX, prim_score = -DBL_MAX;
    int max_idx = -1, prim_idx = -1;
    for (int i = 0; i < n; i++) {
        if ((flag[i] & 0x100) == 0) {
            prim_idx = i;
            prim_score = score[i];
        } else if (max_score < score[i]) {
            max_idx = i;
            max_score = score[i];
        }
    }
    int ties[SP_MAX_ALN_PER_GROUP + 1], nt = 0;
    for (int i = 0; i < n; i++)
        if ((flag[i] & 0x100) != 0 && max_score <= score[i]) ties[nt++] = i;
    if (nt > 1) max_idx = ties[rand() % nt];
    int rnd = rand() % 2;
    double diff = max_score - prim_score;
    int di = (diff >= 2147483648.0 || diff < -2147483648.0 || diff != diff) ? (int) 0x80000000 : (int) diff;
    int ad = di < 0 ? (int) (0u - (unsigned) di) : di;
    if ((double) ad < prim_margin_random) return rnd == 0 ? prim_idx : max_idx;
    if (prim_idx == -1 || max_score <= (prim_score + prim_margin) || max_score < min_score) return prim_idx;
    return max_idx;
}

void merge_and_save_blocks(sph_blocks *t, const char *info_str, const char *bed_path) {  // secphase.c:59-72
    sph_blocks_merge_v2(t);
    fprintf(stderr, "[%s] Total length of %s: %d.\n", get_timestamp(), info_str, (int) sph_blocks_total_length(t));
    fprintf(stderr, "[%s] Total number of %s: %d.\n", get_timestamp(), info_str, (int) sph_blocks_total_number(t));
    if (sph_blocks_save_bed(t, bed_path) != SPH_OK)
        fprintf(stderr, "[%s] Error: Failed to open file %s.\n", get_timestamp(), bed_path);
    fprintf(stderr, "[%s] %s are saved in %s.\n", get_timestamp(), info_str, bed_path);
}

}  // namespace

int main(int argc, char *argv[]) {
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);  // slots x streams per GPU context, see sp_create
    sp_params par;
    sp_params_default(&par);
    int min_var_margin = 50, min_gq = 10;
    (void) min_var_margin;
    (void) min_gq;
    std::string inputPath, fastaPath, variantBedPath, vcfPath, prefix = "secphase", dirPath = "secphase_out_dir";
    bool preset_ont = false, preset_hifi = false, marker_mode = true, write_bam = false;
    int threads = 4, n_gpus = 1, batch_groups = 4096;
    int64_t batch_mb = 256;
    const char *program = strrchr(argv[0], '/');
    program = program ? program + 1 : argv[0];
    int c;
    while (~(c = getopt_long(argc, argv, "i:p:P:G:o:f:v:qd:e:b:n:r:m:ct:s:B:g:@:wxyMh", long_options, nullptr))) {
        switch (c) {
            case 'i': inputPath = optarg; break;
            case 'f': fastaPath = optarg; break;
            case '@': threads = atoi(optarg); break;
            case 'w': write_bam = true; break;
            case 'v': vcfPath = optarg; break;
            case 'P': prefix = optarg; break;
            case 'G': min_gq = atoi(optarg); break;
            case 'o': dirPath = optarg; break;
            case 'x':
                preset_hifi = true;
                sp_params_preset(&par, "hifi");
                break;
            case 'y':
                preset_ont = true;
                sp_params_preset(&par, "ont");
                break;
            case 'q': par.baq_flag = 1; break;
            case 'd': par.conf_d = atof(optarg); break;
            case 'e': par.conf_e = atof(optarg); break;
            case 'b': par.conf_b = atof(optarg); break;
            case 'c': par.consensus = 1; break;
            case 't': par.indel_threshold = atoi(optarg); break;
            case 's': par.set_q = atoi(optarg); break;
            case 'm': par.min_q = atoi(optarg); break;
            case 'p': par.prim_margin_score = atof(optarg); break;
            case 'r': par.prim_margin_random = atof(optarg); break;
            case 'n': par.min_score = atoi(optarg); break;
            case 'g': min_var_margin = atoi(optarg); break;
            case 'B': variantBedPath = optarg; break;
            case 'M': marker_mode = false; break;
            case 'F': par.flank_margin = atoi(optarg); break;
            case 1000: n_gpus = atoi(optarg); break;
            case 1001: batch_groups = atoi(optarg); break;
            case 1002: batch_mb = atoll(optarg); break;
            case 1003:
                printf("%s\n", sp_version());
                return 0;
            default:
                if (c != 'h') fprintf(stderr, "[E::%s] undefined option %c\n", __func__, c);
                usage(program);
                return 1;
        }
    }
    if (inputPath.empty() || fastaPath.empty()) {
        fprintf(stderr, "[E::%s] both -i <INPUT_BAM> and -f <FASTA> are required\n", __func__);
        usage(program);
        return 1;
    }
    if (!vcfPath.empty()) {
        fprintf(stderr, "[%s] Error: variant mode (-v/--inputVcf) is not part of the B200 build; run without -v for marker mode.\n",
                get_timestamp());
        return 1;
    }
    // full-BAQ mode keeps ~700 B of HBM per window row while a batch is in flight (include/secphase_b200.h)
    if (write_bam && batch_groups > 512) batch_groups = 512;
    if (threads < 1) threads = 1;
    if (n_gpus < 1) n_gpus = 1;
    if (batch_groups < 1) batch_groups = 1;
    if (batch_mb < 1) batch_mb = 1;

    struct stat st;
    memset(&st, 0, sizeof(st));
    if (stat(dirPath.c_str(), &st) == -1) mkdir(dirPath.c_str(), 0777);

    auto t_start = std::chrono::steady_clock::now();
    sph_fasta *fa = sph_fasta_load(fastaPath.c_str(), threads);
    if (!fa) {
        fprintf(stderr, "[%s] Error: %s\n", get_timestamp(), sph_last_error());
        return 1;
    }
    // no VCF: an empty table, saved as an empty BED (secphase.c:617-630)
    {
        std::string p = dirPath + "/" + prefix + ".initial_variant_blocks.bed";
        FILE *fp = fopen(p.c_str(), "w");
        if (!fp) fprintf(stderr, "[%s] Failed to open file %s.\n", get_timestamp(), p.c_str());
        else fclose(fp);
    }
    if (preset_ont && preset_hifi) {
        fprintf(stderr, "[%s] Presets --hifi and --ont cannot be enabled at the same time. Select only one of them!\n",
                get_timestamp());
        return EXIT_FAILURE;
    }
    std::string output_log_path = dirPath + "/" + prefix + ".out.log";
    FILE *output_log_file = fopen(output_log_path.c_str(), "w+");
    if (!output_log_file) {
        fprintf(stderr, "[%s] Error: cannot create %s\n", get_timestamp(), output_log_path.c_str());
        return 1;
    }

    sph_bam *bam = sph_bam_open(inputPath.c_str(), threads);
    if (!bam) {
        fprintf(stderr, "[%s] Error: %s\n", get_timestamp(), sph_last_error());
        return 1;
    }
    // contig names link the BAM header and the FASTA, not their order (sam_hdr_tid2name +
    // fai_fetch by name, ptMarker.c:739,819-820): build the codes in BAM tid order
    const int32_t n_tid = sph_bam_n_targets(bam);
    std::vector<int64_t> contig_off((size_t) n_tid + 1, 0);
    std::vector<uint8_t> codes_tid;
    const uint8_t *codes = sph_fasta_codes(fa);
    {
        std::map<std::string, int32_t> by_name;
        for (int32_t i = 0; i < sph_fasta_n(fa); i++) by_name.emplace(sph_fasta_name(fa, i), i);
        bool identity = n_tid == sph_fasta_n(fa);
        std::vector<int32_t> fidx((size_t) n_tid, -1);
        for (int32_t t = 0; t < n_tid; t++) {
            auto it = by_name.find(sph_bam_target_name(bam, t));
            if (it != by_name.end()) fidx[(size_t) t] = it->second;
            if (fidx[(size_t) t] != t) identity = false;
            contig_off[(size_t) t + 1] = contig_off[(size_t) t] + (fidx[(size_t) t] >= 0 ? sph_fasta_len(fa, fidx[(size_t) t]) : 0);
        }
        if (!identity) {
            codes_tid.resize((size_t) contig_off[(size_t) n_tid] + 16);
            const int64_t *fo = sph_fasta_offsets(fa);
            for (int32_t t = 0; t < n_tid; t++)
                if (fidx[(size_t) t] >= 0)
                    memcpy(codes_tid.data() + contig_off[(size_t) t], codes + fo[fidx[(size_t) t]],
                           (size_t) sph_fasta_len(fa, fidx[(size_t) t]));
            codes = codes_tid.data();
            sph_fasta_free(fa);
            fa = nullptr;
        }
        for (int32_t t = 0; t < n_tid; t++)
            if (fidx[(size_t) t] < 0)
                fprintf(stderr, "[%s] Warning: contig %s of the BAM header is not in %s; alignments to it cannot be scored.\n",
                        get_timestamp(), sph_bam_target_name(bam, t), fastaPath.c_str());
    }
    if (n_tid < 1) {
        fprintf(stderr, "[%s] Error: the BAM header lists no reference sequences.\n", get_timestamp());
        return 1;
    }
    {
        std::vector<int64_t> lim((size_t) n_tid);
        for (int32_t t = 0; t < n_tid; t++) lim[(size_t) t] = contig_off[(size_t) t + 1] - contig_off[(size_t) t];
        sph_bam_set_contig_limits(bam, lim.data());
    }

    // one context per device (there is no CPU path: sp_create fails without a GPU).  SECPHASE_B200_DEVICES
    // ("0,0", "2,3", ...) maps context k to a CUDA device: lets --gpus N be exercised on fewer physical GPUs
    // (two contexts on one device behave exactly like two devices to everything above the C ABI).
    std::vector<int> dev_of((size_t) n_gpus);
    for (int d = 0; d < n_gpus; d++) dev_of[(size_t) d] = d;
    if (const char *e = getenv("SECPHASE_B200_DEVICES")) {
        int k = 0;
        for (const char *q = e; *q && k < n_gpus;) {
            dev_of[(size_t) k++] = atoi(q);
            while (*q && *q != ',') q++;
            if (*q == ',') q++;
        }
    }
    // test hook: SECPHASE_B200_FAIL_AT="<ctx>,<n>" makes GPU thread <ctx> report a failure instead of
    // submitting its n-th batch (the error path of a multi-device run must end with exit code 1, not hang)
    int fail_ctx = -1, fail_at = -1;
    if (const char *e = getenv("SECPHASE_B200_FAIL_AT")) sscanf(e, "%d,%d", &fail_ctx, &fail_at);
    std::vector<sp_ctx *> ctx((size_t) n_gpus, nullptr);
    for (int d = 0; d < n_gpus; d++) {
        ctx[(size_t) d] = sp_create(&par, dev_of[(size_t) d]);
        if (!ctx[(size_t) d]) {
            fprintf(stderr, "[%s] Error: cannot initialise GPU %d: %s\n", get_timestamp(), d, sp_last_error());
            return 1;
        }
        if (sp_set_reference_codes(ctx[(size_t) d], n_tid, codes, contig_off.data()) != SP_OK) {
            fprintf(stderr, "[%s] Error: cannot load the assembly on GPU %d: %s\n", get_timestamp(), d, sp_last_error());
            return 1;
        }
        if (write_bam && sp_set_write_qual(ctx[(size_t) d], 1) != SP_OK) {
            fprintf(stderr, "[%s] Error: GPU %d: %s\n", get_timestamp(), d, sp_last_error());
            return 1;
        }
    }
    if (fa) sph_fasta_free(fa);
    fa = nullptr;
    codes_tid.clear();
    codes_tid.shrink_to_fit();
    auto t_ready = std::chrono::steady_clock::now();

    sph_blocks *modified_blocks_by_vars = sph_blocks_create(1);
    sph_blocks *modified_blocks_by_marker = sph_blocks_create(1);
    sph_blocks *variant_blocks_all_haps = sph_blocks_create(0);
    sph_blocks *marker_blocks_all_haps = sph_blocks_create(0);
    int reads_modified_by_vars = 0, reads_modified_by_marker = 0;

    // ---------------------------------------------------------------- pipeline
    Shared shared;
    Queue<sph_batch *> free_q;
    Queue<Work> work_q;
    Queue<Done *> done_q;
    shared.on_fail = [&] {
        work_q.close();
        free_q.close();
    };
    const int n_batches = n_gpus * (SP_N_SLOTS + 1) + 2;
    std::vector<sph_batch *> all_batches;
    for (int i = 0; i < n_batches; i++) {
        sph_batch *hb = sph_batch_create(sp_host_alloc, sp_host_free);
        if (write_bam) sph_batch_keep_records(hb, 1);
        all_batches.push_back(hb);
        free_q.push(hb);
    }
    sph_samw *bam_fo = nullptr;  // secphase.c:643-657
    if (write_bam) {
        std::string p = dirPath + "/" + prefix + ".quality_modified.out.bam";
        bam_fo = sph_samw_open_mt(p.c_str(), bam, threads);
        if (!bam_fo) {
            fprintf(stderr, "[%s] Error: %s\n", get_timestamp(), sph_last_error());
            return 1;
        }
    }
    fprintf(stderr, "[%s] Started parsing alignments\n", get_timestamp());

    std::atomic<int64_t> ingest_us{0}, submit_us{0}, wait_us{0}, starved_us{0};
    auto now_us = [] {
        return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
    };
    std::thread reader([&] {
        int64_t seq = 0;
        for (;;) {
            sph_batch *hb = nullptr;
            if (!free_q.pop(hb)) break;
            if (shared.is_failed()) break;
            int64_t t0 = now_us();
            int32_t n = sph_bam_next_batch(bam, hb, batch_groups, batch_mb << 20);
            ingest_us += now_us() - t0;
            if (n < 0) {
                shared.fail(sph_last_error());
                break;
            }
            if (n == 0) break;
            work_q.push(Work{hb, seq++});
        }
        work_q.close();
    });

    std::vector<std::thread> gpu_threads;
    std::mutex gpu_exit_mu;
    int gpu_live = n_gpus;
    for (int d = 0; d < n_gpus; d++) {
        gpu_threads.emplace_back([&, d] {
            sp_ctx *cx = ctx[(size_t) d];
            struct InFlight {
                Work w;
                int slot;
            };
            std::deque<InFlight> inflight;
            int next_slot = 0;
            auto collect = [&]() -> bool {
                InFlight f = inflight.front();
                inflight.pop_front();
                sp_result r;
                int64_t t0 = now_us();
                int wrc = sp_wait(cx, f.slot, &r);
                wait_us += now_us() - t0;
                if (wrc != SP_OK) {
                    shared.fail(std::string("GPU ") + std::to_string(d) + ": " + sp_last_error());
                    return false;
                }
                Done *dn = new Done();
                dn->hb = f.w.hb;
                dn->seq = f.w.seq;
                dn->group.assign(r.group, r.group + (size_t) r.n_groups * SP_GROUP_W);
                dn->score.assign(r.score, r.score + r.n_alns);
                dn->extent.assign(r.extent, r.extent + (size_t) r.n_alns * 4);
                dn->marker_off.assign(r.marker_off, r.marker_off + r.n_groups + 1);
                dn->marker.assign(r.marker, r.marker + (size_t) r.marker_off[r.n_groups] * SP_MARKER_W);
                if (r.baq_qual) dn->baq_qual.assign(r.baq_qual, r.baq_qual + r.baq_qual_bytes);
                dn->hmm_instances = r.hmm_instances;
                dn->hmm_cells = r.hmm_cells;
                dn->gpu_ms = r.ms_total;
                dn->hmm_ms = r.ms_hmm;
                dn->launches = r.gpu_launches;
                done_q.push(dn);
                return true;
            };
            Work w;
            bool ok = true;
            int n_submitted = 0;
            for (;;) {
                if (shared.is_failed()) break;
                // with batches in flight never block on the reader: retire finished ones meanwhile
                int64_t tq = now_us();
                int st = inflight.empty() ? (work_q.pop(w) ? 1 : -1) : work_q.try_pop(w);
                if (inflight.empty()) starved_us += now_us() - tq;
                if (st < 0) break;  // closed and drained
                if (st == 0) {
                    int done = sp_poll(cx, inflight.front().slot);
                    if (done < 0) {
                        shared.fail(std::string("GPU ") + std::to_string(d) + ": " + sp_last_error());
                        ok = false;
                        break;
                    }
                    if (done == 1) {
                        if (!(ok = collect())) break;
                    } else {
                        std::this_thread::sleep_for(std::chrono::microseconds(50));
                    }
                    continue;
                }
                if ((int) inflight.size() == SP_N_SLOTS && !(ok = collect())) break;
                if (d == fail_ctx && n_submitted++ == fail_at) {
                    shared.fail(std::string("GPU ") + std::to_string(d) + ": injected failure (SECPHASE_B200_FAIL_AT)");
                    ok = false;
                    break;
                }
                int64_t t0 = now_us();
                int src = sp_submit(cx, sph_batch_view(w.hb), next_slot);
                submit_us += now_us() - t0;
                if (src != SP_OK) {
                    shared.fail(std::string("GPU ") + std::to_string(d) + ": " + sp_last_error());
                    ok = false;
                    break;
                }
                inflight.push_back({w, next_slot});
                next_slot = (next_slot + 1) % SP_N_SLOTS;
            }
            while (ok && !inflight.empty()) ok = collect();
            std::lock_guard<std::mutex> g(gpu_exit_mu);
            if (--gpu_live == 0) done_q.close();
        });
    }

    // ---------------------------------------------------------------- ordered output (this thread)
    int64_t next_seq = 0, total_groups = 0, total_alns = 0, hmm_instances = 0, hmm_cells = 0, launches = 0;
    double gpu_ms = 0, hmm_ms = 0;
    std::map<int64_t, Done *> parked;
    int64_t alignment_log_idx = 1;
    const int64_t alignment_log_size = 20000;  // secphase.c:249
    std::string rec;
    auto emit = [&](Done *dn) {
        const sp_flat_batch *b = sph_batch_view(dn->hb);
        const int G = b->n_groups;
        // secphase.c:182-189: every record of every job of the marker branch, qualities as BAQ left them
        if (bam_fo && marker_mode && b->n_alns > 0) {
            if (sph_samw_write_batch(bam_fo, dn->hb, dn->baq_qual.empty() ? nullptr : dn->baq_qual.data()) != SPH_OK)
                shared.fail(std::string("writing the output bam: ") + sph_last_error());
        }
        for (int g = 0; g < G; g++) {
            const int a0 = b->grp_aln_off[g], n = b->grp_aln_off[g + 1] - a0;
            const int32_t *row = &dn->group[(size_t) g * SP_GROUP_W];
            int best = row[0];
            if (n_gpus > 1)
                best = select_best_host(n, b->flag + a0, dn->score.data() + a0, par.prim_margin_score,
                                        (double) par.min_score, par.prim_margin_random);
            if (!marker_mode) continue;  // -M without a VCF: nothing is scored (secphase.c:156)
            if (best < 0 || !(b->flag[a0 + best] & 0x100)) continue;
            // secphase.c:194-216
            const char *contig[SP_MAX_ALN_PER_GROUP];
            int32_t rfe[SP_MAX_ALN_PER_GROUP];
            for (int i = 0; i < n; i++) {
                contig[i] = sph_bam_target_name(bam, b->tid[a0 + i]);
                rfe[i] = dn->extent[(size_t) (a0 + i) * 4 + 1];
            }
            const char *qn = b->qname_pool + b->qname_off[g];
            int32_t qn_len = (int32_t) (b->qname_off[g + 1] - b->qname_off[g]);
            int64_t need = sph_format_marker_record(nullptr, 0, qn, qn_len, n, b->flag + a0, dn->score.data() + a0,
                                                    contig, b->pos + a0, rfe, best);
            rec.resize((size_t) need);
            sph_format_marker_record(&rec[0], need, qn, qn_len, n, b->flag + a0, dn->score.data() + a0, contig,
                                     b->pos + a0, rfe, best);
            fwrite(rec.data(), 1, rec.size(), output_log_file);
            fflush(output_log_file);
            const int prim = row[1];
            for (int idx : {prim, best}) {
                if (idx < 0) continue;
                const int32_t *e = &dn->extent[(size_t) (a0 + idx) * 4];
                sph_blocks_add(modified_blocks_by_marker, contig[idx], e[0], e[1]);
                for (int64_t m = dn->marker_off[(size_t) g]; m < dn->marker_off[(size_t) g + 1]; m++) {
                    const int32_t *mk = &dn->marker[(size_t) m * SP_MARKER_W];
                    if (mk[0] == idx) sph_blocks_add(marker_blocks_all_haps, contig[idx], mk[5], mk[5]);
                }
            }
            reads_modified_by_marker++;
        }
        total_groups += G;
        total_alns += b->n_alns;
    };
    Done *dn = nullptr;
    while (done_q.pop(dn)) {
        parked[dn->seq] = dn;
        // after a failure the batch a dead GPU thread held never arrives: nothing more is emitted, later
        // batches are dropped as they come (their buffers are freed with all_batches below)
        while (!parked.empty() && (parked.begin()->first == next_seq || shared.is_failed())) {
            Done *d2 = parked.begin()->second;
            parked.erase(parked.begin());
            if (!shared.is_failed()) emit(d2);
            hmm_instances += d2->hmm_instances;
            hmm_cells += d2->hmm_cells;
            gpu_ms += d2->gpu_ms;
            hmm_ms += d2->hmm_ms;
            launches += d2->launches;
            free_q.push(d2->hb);
            delete d2;
            next_seq++;
            int64_t pa = 0, pr = 0;
            sph_bam_counts(bam, &pa, &pr);
            if (pa >= alignment_log_idx * alignment_log_size) {  // secphase.c:320-334
                alignment_log_idx = pa / alignment_log_size + 1;
                fprintf(stderr,
                        "[%s] #parsed alignments = %lld, #parsed reads = %lld, #modifed by phased variants = %d, #modifed by markers = %d\n",
                        get_timestamp(), (long long) pa, (long long) pr, reads_modified_by_vars, reads_modified_by_marker);
                fflush(stderr);
            }
        }
    }
    free_q.close();
    reader.join();
    for (auto &t : gpu_threads) t.join();
    for (auto &kv : parked) delete kv.second;
    if (bam_fo && sph_samw_close(bam_fo) != SPH_OK) shared.fail("closing the output bam failed");
    if (shared.is_failed()) {
        fprintf(stderr, "[%s] Error: %s\n", get_timestamp(), shared.error.c_str());
        fclose(output_log_file);
        return 1;
    }
    auto t_scored = std::chrono::steady_clock::now();
    if (sph_bam_skipped_groups(bam) > 0)
        fprintf(stderr, "[%s] Warning: %lld read groups skipped (CIGAR N/P operations, or SEQ/QUAL absent or inconsistent with the CIGAR).\n",
                get_timestamp(), (long long) sph_bam_skipped_groups(bam));

    fprintf(stderr, "[%s] Number of reads modified by phased variants = %d\n", get_timestamp(), reads_modified_by_vars);
    fprintf(stderr, "[%s] Number of reads modified by marker score = %d\n", get_timestamp(), reads_modified_by_marker);

    std::string bed;
    bed = dirPath + "/" + prefix + ".modified_read_blocks.variants.bed";
    merge_and_save_blocks(modified_blocks_by_vars, "read blocks modified by phased variants", bed.c_str());
    bed = dirPath + "/" + prefix + ".modified_read_blocks.markers.bed";
    merge_and_save_blocks(modified_blocks_by_marker, "read blocks modified by markers", bed.c_str());
    bed = dirPath + "/" + prefix + ".variant_blocks.bed";
    merge_and_save_blocks(variant_blocks_all_haps, "projected variant blocks on all haplotypes", bed.c_str());
    bed = dirPath + "/" + prefix + ".marker_blocks.bed";
    merge_and_save_blocks(marker_blocks_all_haps, "projected marker blocks on all haplotypes", bed.c_str());
    fclose(output_log_file);

    // one machine-readable summary line for benchmarking (stderr, after the reference's lines)
    {
        auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
            return std::chrono::duration<double>(b - a).count();
        };
        int64_t pa = 0, pr = 0;
        sph_bam_counts(bam, &pa, &pr);
        fprintf(stderr,
                "[secphase_b200] {\"read_groups\": %lld, \"alignments\": %lld, \"parsed_alignments\": %lld, \"parsed_reads\": %lld, "
                "\"hmm_instances\": %lld, \"hmm_cells\": %lld, \"gpus\": %d, \"host_threads\": %d, \"setup_s\": %.3f, "
                "\"score_s\": %.3f, \"gpu_busy_ms\": %.1f, \"hmm_ms\": %.1f, \"gpu_launches\": %lld, \"ingest_s\": %.3f, "
                "\"submit_s\": %.3f, \"wait_s\": %.3f, \"gpu_starved_s\": %.3f, \"total_s\": %.3f}\n",
                (long long) total_groups, (long long) total_alns, (long long) pa, (long long) pr, (long long) hmm_instances,
                (long long) hmm_cells, n_gpus, threads, secs(t_start, t_ready), secs(t_ready, t_scored), gpu_ms, hmm_ms,
                (long long) launches, ingest_us.load() * 1e-6, submit_us.load() * 1e-6, wait_us.load() * 1e-6,
                starved_us.load() * 1e-6, secs(t_start, std::chrono::steady_clock::now()));
    }
    for (sph_batch *hb : all_batches) sph_batch_destroy(hb);
    sph_blocks_destroy(modified_blocks_by_vars);
    sph_blocks_destroy(modified_blocks_by_marker);
    sph_blocks_destroy(variant_blocks_all_haps);
    sph_blocks_destroy(marker_blocks_all_haps);
    sph_bam_close(bam);
    for (sp_ctx *cx : ctx) sp_destroy(cx);
    return 0;
}
