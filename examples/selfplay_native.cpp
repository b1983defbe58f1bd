// A native host for libothello_b200 (no Python, no PyTorch): allocates every buffer the engine asks
// for with cudaMalloc, plays complete self-play games with the device hash-stub evaluator through
// the C ABI (include/othello_b200.h), drains the replay tuples and prints throughput.
//
//   nvcc -O2 -o examples/selfplay_native examples/selfplay_native.cpp \
//        -Lalphazero_othello_b200 -lothello_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../alphazero_othello_b200'
//   ./examples/selfplay_native [n_slots=4096] [sims=100] [games_per_slot=1]
//
// With a real network the loop body becomes: run the net on `nn_input` (cuDNN / TensorRT), then
// oth_mcts_step(&cfg, &bufs, priors, values, nn_input, stream) -- see INTEGRATION.md.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../include/othello_b200.h"

#define CHECK(x)                                                                                   \
    do {                                                                                           \
        int rc_ = (x);                                                                             \
        if (rc_ != OTH_OK) {                                                                       \
            fprintf(stderr, "%s -> %s [%s]\n", #x, oth_error_string(rc_), oth_last_cuda_error()); \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

int main(int argc, char** argv)
{
    if (oth_device_count() <= 0) {
        fprintf(stderr, "no CUDA device: libothello_b200 has no CPU fallback\n");
        return 2;
    }
    oth_mcts_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.n_slots = argc > 1 ? atoi(argv[1]) : 4096;
    cfg.num_simulations = argc > 2 ? atoi(argv[2]) : 100;
    cfg.games_per_slot = argc > 3 ? atoi(argv[3]) : 1;
    cfg.node_cap = 40 * cfg.num_simulations + 1024;
    cfg.path_cap = 128;
    cfg.num_exploratory_moves = 35;
    cfg.eval_kind = OTH_EVAL_STUB_H;
    cfg.self_play = 1;
    cfg.max_inline_sims = 16;
    cfg.lanes = 8;
    cfg.split_stub = 1;   // the production kernel pair (hot step kernel + move kernel) with the stub in the network's place
    cfg.move_launch = 1;  // the step kernel launches the move kernel from the device when a move is due
    cfg.out_game_cap = (int64_t)cfg.n_slots * cfg.games_per_slot + 16;
    cfg.out_pos_cap = cfg.out_game_cap * 72;
    cfg.c_puct = 2.0;
    cfg.dirichlet_alpha = 1.0;
    cfg.dirichlet_epsilon = 0.3;
    cfg.temperature = 1.0;
    cfg.lambda = 0.98;
    cfg.seed = 42;
    cfg.game_id_base = 0;
    cfg.game_id_stride = cfg.n_slots;
    cfg.stub_salt = 7;

    int64_t bytes[OTH_BUF_COUNT];
    CHECK(oth_mcts_buffer_bytes(&cfg, bytes));
    oth_mcts_buffers bufs;
    memset(&bufs, 0, sizeof(bufs));  // .profile = NULL: no launch timing
    size_t total = 0;
    for (int i = 0; i < OTH_BUF_COUNT; i++) {
        const size_t nb = bytes[i] > 0 ? (size_t)bytes[i] : 256;
        if (cudaMalloc(&bufs.buf[i], nb) != cudaSuccess || cudaMemset(bufs.buf[i], 0, nb) != cudaSuccess) {
            fprintf(stderr, "cudaMalloc of buffer %d (%zu bytes) failed\n", i, nb);
            return 1;
        }
        total += nb;
    }
    float* nn_input = nullptr;  // unused by the stub evaluator but part of the call
    cudaMalloc(&nn_input, (size_t)cfg.n_slots * 64 * sizeof(float));
    cudaStream_t stream;
    cudaStreamCreate(&stream);

    CHECK(oth_mcts_reset(&cfg, &bufs, stream));
    std::vector<unsigned long long> counters(16);
    const auto t0 = std::chrono::steady_clock::now();
    long launches = 0;
    for (;;) {
        for (int i = 0; i < 64; i++, launches++) CHECK(oth_mcts_step(&cfg, &bufs, nullptr, nullptr, nn_input, stream));
        CHECK(oth_mcts_poll(&cfg, &bufs, stream));
        cudaMemcpyAsync(counters.data(), bufs.buf[OTH_BUF_COUNTERS], 16 * 8, cudaMemcpyDeviceToHost, stream);
        cudaStreamSynchronize(stream);
        if (counters[OTH_CNT_ERRORS]) {
            fprintf(stderr, "%llu slot(s) in error (arena / ring overflow?)\n", counters[OTH_CNT_ERRORS]);
            return 1;
        }
        if (counters[OTH_CNT_ACTIVE] == 0) break;
    }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    // drain: replay tuples (board, pi, value target) of every finished game
    const long n_pos = (long)counters[OTH_CNT_POSITIONS], n_games = (long)counters[OTH_CNT_OUT_GAMES];
    std::vector<double> values(n_pos);
    std::vector<long long> games((size_t)n_games * 4);
    cudaMemcpy(values.data(), bufs.buf[OTH_BUF_OUT_VALUE], n_pos * sizeof(double), cudaMemcpyDeviceToHost);
    cudaMemcpy(games.data(), bufs.buf[OTH_BUF_OUT_GAMES], games.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    long wins[3] = {0, 0, 0};
    for (long g = 0; g < n_games; g++) wins[games[4 * g + 3] + 1]++;
    double vsum = 0;
    for (double v : values) vsum += v;
    printf("{\"slots\": %d, \"sims_per_move\": %d, \"games\": %ld, \"positions\": %ld, \"launches\": %ld, \"seconds\": %.3f, "
           "\"sims\": %llu, \"sims_per_s\": %.0f, \"white_wins\": %ld, \"draws\": %ld, \"black_wins\": %ld, \"mean_value_target\": %.4f, "
           "\"device_bytes\": %zu}\n",
           cfg.n_slots, cfg.num_simulations, n_games, n_pos, launches, sec, counters[OTH_CNT_SIMS], counters[OTH_CNT_SIMS] / sec, wins[0],
           wins[1], wins[2], n_pos ? vsum / n_pos : 0.0, total);
    for (int i = 0; i < OTH_BUF_COUNT; i++) cudaFree(bufs.buf[i]);
    cudaFree(nn_input);
    return n_games == (long)cfg.n_slots * cfg.games_per_slot ? 0 : 1;
}
