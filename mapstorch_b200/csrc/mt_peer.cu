// Cross-GPU completion signalling for the element maps that the fit kernels store into other GPUs'
// buffers over NVLink (dist.PeerMaps).
//
// The exchange itself is fused into K3's epilogue; what remains of the "collective" is to tell every
// destination GPU that this rank's stores of step k have all landed.  Two tiny kernels on the fit's stream:
//   mt_peer_signal : ++epoch (device counter), then epoch -> flags[r][my rank] on every rank r.  Stream order puts
//                    it after the fit kernels; kernel completion makes their peer stores visible system-wide, the
//                    flag stores are release stores at system scope on top of that.
//   mt_peer_wait   : spin (acquire loads, bounded) until every rank's slot in MY flag array has reached my epoch - lag.
// Both read the epoch from device memory, so a CUDA graph holding them replays correctly.  Because signal and wait
// are separate and the wait takes a lag, step k can end with "wait for everybody's step k-1" (double-buffered maps):
// the ranks then synchronise with one step of slack instead of in lock-step, and the barrier latency hides behind
// the next step's streaming.
#include "mt_common.cuh"

namespace {

__global__ void peer_signal_kernel(MtPeerSync sync, uint32_t *epoch) {
    __shared__ uint32_t next;
    if (threadIdx.x == 0) {
        next = *epoch + 1u;
        *epoch = next;
    }
    __syncthreads();
    if ((int)threadIdx.x < sync.n_ranks) {
        __threadfence_system();
        uint32_t *slot = sync.flags[threadIdx.x] + sync.rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(slot), "r"(next) : "memory");
    }
}

__global__ void peer_wait_kernel(MtPeerSync sync, const uint32_t *epoch, uint32_t lag, int *timed_out) {
    if ((int)threadIdx.x < sync.n_ranks) {
        const uint32_t want = *epoch - lag;
        const uint32_t *slot = sync.flags[sync.rank] + threadIdx.x;
        long long t0 = 0;
        for (uint32_t spin = 0;; ++spin) {
            uint32_t seen;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(slot) : "memory");
            if ((int32_t)(seen - want) >= 0) break;
            if ((spin & 0x3FFu) == 0x3FFu) {
                // bounded: a rank that died must surface as an error on the others, not hang their GPUs
                const long long now = clock64();
                if (t0 == 0) t0 = now;
                if (now - t0 > 20000000000ll) {
                    if (timed_out) atomicExch(timed_out, 1 + (int)threadIdx.x);
                    break;
                }
            }
        }
    }
}

int check_sync(const MtPeerSync *s, const char *who) {
    MT_REQUIRE(s, "%s: null MtPeerSync", who);
    MT_REQUIRE(s->n_ranks >= 1 && s->n_ranks <= MT_MAX_PEERS && s->rank >= 0 && s->rank < s->n_ranks,
               "%s: n_ranks=%d rank=%d", who, s->n_ranks, s->rank);
    for (int r = 0; r < s->n_ranks; ++r) MT_REQUIRE(s->flags[r], "%s: flags[%d] is null", who, r);
    return MT_OK;
}

}  // namespace

extern "C" int mt_peer_signal(const MtPeerSync *sync_host, uint32_t *epoch_dev, void *stream) {
    if (int rc = mt::require_sm100()) return rc;
    if (int rc = check_sync(sync_host, "mt_peer_signal")) return rc;
    MT_REQUIRE(epoch_dev, "mt_peer_signal: null epoch");
    peer_signal_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(*sync_host, epoch_dev);
    MT_CUDA_TRY(cudaGetLastError());
    return MT_OK;
}

extern "C" int mt_peer_wait(const MtPeerSync *sync_host, const uint32_t *epoch_dev, int32_t lag, int32_t *timed_out_dev,
                            void *stream) {
    if (int rc = mt::require_sm100()) return rc;
    if (int rc = check_sync(sync_host, "mt_peer_wait")) return rc;
    MT_REQUIRE(epoch_dev && lag >= 0, "mt_peer_wait: null epoch or negative lag");
    peer_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(*sync_host, epoch_dev, (uint32_t)lag, timed_out_dev);
    MT_CUDA_TRY(cudaGetLastError());
    return MT_OK;
}
