"""CPU restatement of the sampler tail (TEST INFRASTRUCTURE, see oracle/__init__.py).

``generate_step_with_noise`` follows /root/reference/src/pgen/esm_sampler.py:23-45 line by line, except that
the Exp(1) variates ``Categorical.sample()`` would draw from torch's generator are passed in, using the
identity  Categorical(logits=v).sample() == argmax(softmax(v - logsumexp(v)) / q),  q ~ Exp(1)
(torch.multinomial's n_sample == 1 path; pinned against the real ``torch.distributions`` call in
tests/test_oracle.py and against the reference's own ``generate_step`` in tests/golden/).
"""
import torch


def effective_k(top_k, n, sample):
    # esm_sampler.py:32-38
    return n if (sample or top_k <= 0 or top_k > n) else top_k


def generate_step_with_noise(logits_row, noise_row, valid_idx, top_k=0, temperature=None, sample=False):
    logits = logits_row
    if temperature is not None:
        logits = logits / temperature                       # :24-25
    sub = logits[valid_idx]                                 # :30
    k = effective_k(top_k, len(sub), sample)
    vals, idx = sub.topk(k)                                 # :40
    norm = vals - vals.logsumexp(dim=-1, keepdim=True)      # Categorical.__init__(logits=...)
    probs = torch.softmax(norm, dim=-1)                     # Categorical.probs
    draw = int(torch.argmax(probs / noise_row[:k]))         # multinomial(probs, 1): argmax(p / q)
    return int(valid_idx[int(idx[draw])])                   # :43-45
