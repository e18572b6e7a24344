// shard.cu -- sharded state over 2/4/8 GPUs (placeholder until the exchange kernels land).
#include "common.cuh"

namespace qi {

int shard_prepare_gate(qi_state* s, const qi_gate* g, PhysGate* out, bool* skip) {
    (void)s; (void)g; (void)out; (void)skip;
    return fail(QI_ERR_PEER, 0, 0, "sharded states are not available in this build");
}
bool shard_needs_exchange(const qi_state* s, const qi_gate* g) { (void)s; (void)g; return false; }
int shard_do_exchange(qi_state* s, const qi_gate* g) { (void)s; (void)g; return QI_OK; }
int shard_allreduce_sum(qi_state* s, double* host_vals, int count) { (void)s; (void)host_vals; (void)count; return QI_OK; }
int shard_localise_mask(qi_state* s, const qi_pauli_term* t) { (void)s; (void)t; return QI_OK; }

}  // namespace qi

extern "C" {
void qi_shard_release(qi_state* s) { (void)s; }
int qi_shard_new_zero(uint32_t, int, int, qi_state**) { return qi::fail(QI_ERR_PEER, 0, 0, "not built"); }
int qi_shard_new_plus(uint32_t, int, int, qi_state**) { return qi::fail(QI_ERR_PEER, 0, 0, "not built"); }
int qi_shard_new_basis_n(uint32_t, uint64_t, int, int, qi_state**) { return qi::fail(QI_ERR_PEER, 0, 0, "not built"); }
int qi_shard_export(qi_state*, uint8_t*) { return qi::fail(QI_ERR_PEER, 0, 0, "not built"); }
int qi_shard_attach(qi_state*, const uint8_t*) { return qi::fail(QI_ERR_PEER, 0, 0, "not built"); }
int qi_shard_rank(const qi_state* s) { return s ? s->rank : 0; }
int qi_shard_world(const qi_state* s) { return s ? s->world : 1; }
int qi_shard_comm_stats(const qi_state* s, uint64_t* a, uint64_t* b, uint64_t* c) {
    if (!s) return qi::fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "state is NULL");
    if (a) *a = s->bytes_sent; if (b) *b = s->bytes_recv; if (c) *c = s->exchanges;
    return QI_OK;
}
}
