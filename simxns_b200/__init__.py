"""simxns_b200 — B200-native (sm_100a) dual-encoder training step behind the SimANS / PROD / LEAD model API.

    from simxns_b200.models import BiBertEncoder, HFBertEncoder, Reranker, BiEncoderNllLoss
    from simxns_b200.losses import fused_score_loss, simans_distill_loss

The compute lives in simxns_b200/libsimx_b200.so (C ABI: include/simx_b200.h); build it with
`python -c "import __graft_entry__ as g; g.build()"` or `make -C simxns_b200/csrc`.
"""
__version__ = "0.1.0"
