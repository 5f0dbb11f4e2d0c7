// zk_encode.cu -- batched Zstandard frame compression for sm_100a (stub being filled in)
#include "zk_encode.h"
size_t zk_encode_bound(size_t n, uint32_t frame_size) {
    if (frame_size == 0) frame_size = 1;
    size_t frames = n / frame_size + 1;
    size_t blocks = n / 32768 + frames + 1;
    return n + frames * 32 + blocks * 4 + 64;
}
void zk_encode_ws_free(ZkEncodeWs* ws) {
    if (ws->buf) cudaFree(ws->buf);
    if (ws->h_sizes) cudaFreeHost(ws->h_sizes);
    *ws = ZkEncodeWs();
}
