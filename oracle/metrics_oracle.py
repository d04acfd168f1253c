"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain Python / numpy) of the reference's cluster-quality scores.
Only tests/ may import this; nothing under smartcore_b200/ does.

Follows, line by line:
  contingency_matrix   /root/reference/src/metrics/cluster_helpers.rs:7-25  (unique_with_indices: arrays.rs:233-247)
  entropy              /root/reference/src/metrics/cluster_helpers.rs:27-48
  mutual_info_score    /root/reference/src/metrics/cluster_helpers.rs:50-104
  HCVScore::compute    /root/reference/src/metrics/cluster_hcv.rs:36-55

PARITY PIN: the reference's own known-answer tests -- contingency_matrix_test (cluster_helpers.rs:117-125),
entropy_test 1.2770 (:131-135), mutual_info_score_test 0.3254 (:141-147), homogeneity_score 0.2548 / 0.5440 / 0.3471
(cluster_hcv.rs:94-104) -- are checked in tests/test_oracle.py.  The reference sums entropy terms in HashMap order
(unspecified); this restatement uses ascending label order.
"""
import math


def unique_with_indices(v):
    unique = sorted(set(int(e) for e in v))
    pos = {u: i for i, u in enumerate(unique)}
    return unique, [pos[int(e)] for e in v]


def contingency_matrix(labels_true, labels_pred):
    classes, class_idx = unique_with_indices(labels_true)
    clusters, cluster_idx = unique_with_indices(labels_pred)
    m = [[0] * len(clusters) for _ in classes]
    for i in range(len(class_idx)):
        m[class_idx[i]][cluster_idx[i]] += 1
    return m


def entropy(data):
    bincounts = {}
    for e in data:
        k = int(e)
        bincounts[k] = bincounts.get(k, 0) + 1
    ent = 0.0
    total = sum(bincounts.values())
    for k in sorted(bincounts):
        c = bincounts[k]
        if c > 0:
            pi = float(c)
            ent -= (pi / float(total)) * (math.log(pi) - math.log(float(total)))
    return ent


def mutual_info_score(contingency):
    contingency_sum = 0
    pi = [0] * len(contingency)
    pj = [0] * len(contingency[0])
    nzx, nzy, nz_val = [], [], []
    for r in range(len(contingency)):
        for c in range(len(contingency[0])):
            contingency_sum += contingency[r][c]
            pi[r] += contingency[r][c]
            pj[c] += contingency[r][c]
            if contingency[r][c] > 0:
                nzx.append(r); nzy.append(c); nz_val.append(contingency[r][c])
    csum = float(contingency_sum)
    csum_ln = math.log(csum)
    pi_sum_l = math.log(float(sum(pi)))
    pj_sum_l = math.log(float(sum(pj)))
    result = 0.0
    for i in range(len(nz_val)):
        log_nm = math.log(float(nz_val[i]))
        nm = float(nz_val[i]) / csum
        log_outer = -math.log(float(pi[nzx[i]] * pj[nzy[i]])) + pi_sum_l + pj_sum_l
        result += (nm * (log_nm - csum_ln)) + nm * log_outer
    return max(result, 0.0)


def hcv(y_true, y_pred):
    entropy_c = entropy(y_true)
    entropy_k = entropy(y_pred)
    mi = mutual_info_score(contingency_matrix(y_true, y_pred))
    nan = float("nan")
    h = mi / entropy_c if entropy_c != 0.0 else (nan if mi == 0.0 else math.copysign(math.inf, mi))
    c = mi / entropy_k if entropy_k != 0.0 else (nan if mi == 0.0 else math.copysign(math.inf, mi))
    v = 0.0 if h + c == 0.0 else 2.0 * h * c / (1.0 * h + c)
    return h, c, v
