"""demo2_b200 -- B200-native (sm_100a) retrieval hot path of maxingan2412/DeMo2.

Drop-in modules (same names / signatures as the reference):
    demo2_b200.metrics       <- utils/metrics.py      (R1_mAP_eval, eval_func, euclidean_distance, cosine_similarity)
    demo2_b200.reranking     <- utils/reranking.py    (re_ranking, both call forms)
    demo2_b200.triplet_loss  <- layers/triplet_loss.py (TripletLoss, hard_example_mining, euclidean_dist, ...)
    demo2_b200.parallel      gallery-sharded multi-GPU evaluation (new)
All arithmetic runs in libdemo_b200.so (C ABI, include/demo_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
