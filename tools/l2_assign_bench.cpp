// Host-only timing of the L2 circuit's witness assignment at the scaled shape (64 transfers over 128 accounts, SURVEY.md 8d
// config 1): sequential versus the two-pass parallel walkers.  The CUDA-side symbols zkb_l2_prove needs are stubbed.
//   g++ -O3 -std=c++17 -pthread -o /tmp/l2_assign_bench tools/l2_assign_bench.cpp
//   for t in 1 2 4 8 16; do ZKB_L2_ASSIGN_THREADS=$t /tmp/l2_assign_bench | tail -1; done
#include "../zelana_b200/csrc/l2_circuit.cpp"
#include <chrono>
#include <cstdio>
#include <random>
extern "C" int zkb_prove(zkb_ctx*, const zkb_pk*, const zkb_r1cs*, const uint8_t*, const uint8_t*, const uint8_t*, uint8_t*, uint8_t*, uint8_t*) { return 0; }
extern "C" const char* zkb_last_error(zkb_ctx*) { return ""; }
extern "C" int zkb_ctx_create(int, zkb_ctx**) { return -1; }
extern "C" void zkb_ctx_destroy(zkb_ctx*) {}
extern "C" int zkb_ctx_set_blocking_sync(zkb_ctx*, int) { return 0; }
int main() {
  using clk = std::chrono::steady_clock;
  std::mt19937_64 rng(1);
  const int NA = 128, NT = 64;
  std::vector<uint8_t> pk(NA * 32), snd(NT * 32), rcp(NT * 32);
  for (auto& x : pk) x = (uint8_t)rng();
  std::vector<uint64_t> bal(NA, 1000000000ull), amt(NT);
  for (int i = 0; i < NT; ++i) { memcpy(&snd[32 * i], &pk[32 * (rng() % NA)], 32); memcpy(&rcp[32 * i], &pk[32 * (rng() % NA)], 32); amt[i] = rng() % 1000; }
  zkb_l2_witness w; memset(&w, 0, sizeof w);
  w.account_pks = pk.data(); w.account_balances = bal.data(); w.n_accounts = NA;
  w.tx_senders = snd.data(); w.tx_recipients = rcp.data(); w.tx_amounts = amt.data(); w.n_txs = NT;
  zkb_l2_circuit* c; zkb_l2_circuit_create(&w, &c);
  zkb_l2_public_inputs in; uint8_t z32[32] = {0};
  zkb_l2_roots(&w, 5, z32, &in);
  std::vector<uint8_t> z((8 + c->num_witness) * 32);
  for (int rep = 0; rep < 3; ++rep) {
    auto t0 = clk::now();
    std::vector<l2::Fr> zz;
    int rc = l2_assign(c, &in, &w, &zz);
    auto t1 = clk::now();
    l2_z_to_bytes(zz, z.data());
    auto t2 = clk::now();
    int sat = 0; zkb_l2_circuit_is_satisfied(c, z.data(), &sat, nullptr);
    printf("threads %d rc %d assign %.2f ms bytes %.2f ms satisfied %d\n", l2_assign_threads(), rc,
           std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t2 - t1).count(), sat);
  }
}
