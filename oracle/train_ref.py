"""ORACLE (test infrastructure): CPU restatement of the reference's loss path and greedy decode loop.

Restates, with the config values passed in explicitly instead of read from import-time globals:
  * configs/common/__init__.py:31-57   vocab_size / start_idx arithmetic
  * train.py:20                        length_tensor
  * train.py:79-111                    make_distributions
  * train.py:114-131                   pick_distributions_by_prev_token
  * train.py:133-138                   filtered_logit  (log_softmax over dim=1, i.e. the SEQUENCE axis — F4)
  * train.py:160-165                   one loss evaluation of the hot loop
  * scripts/generate_midi_many.py:13-56  greedy generate() (full re-forward on a sliding window)

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

from collections import Counter

import torch
import torch.nn.functional as F

# configs/common/config.yaml:1-7
DISCRETIZATION = dict(pitch=128, dyn=128, length=512, time=512, channel=129, tempo=250)


def vocab_layout(disc=DISCRETIZATION):
    """configs/common/__init__.py:31-57 -> (vocab_size, start_idx)."""
    vocab_size = sum([disc["pitch"] * disc["channel"], disc["dyn"], disc["length"], disc["time"], disc["tempo"]])
    offset = 0
    start_idx = {}
    start_idx["pitch"] = offset
    offset += disc["pitch"] * disc["channel"]
    start_idx["dyn"] = offset
    offset += disc["dyn"]
    start_idx["length"] = offset
    offset += disc["length"]
    start_idx["time"] = offset
    offset += disc["time"]
    start_idx["tempo"] = offset
    return vocab_size, start_idx


def make_distributions(device="cpu", disc=DISCRETIZATION):
    """train.py:79-111."""
    vocab_size, start_idx = vocab_layout(disc)
    length_tensor = torch.linspace(1, 3, steps=disc["length"] - 1).to(device)  # train.py:20
    distributions = torch.zeros(5, vocab_size, device=device)
    start = [start_idx["pitch"], start_idx["dyn"], start_idx["length"], start_idx["time"], start_idx["tempo"]]
    end = [start_idx["dyn"] - 1, start_idx["length"] - 1, start_idx["time"] - 1, start_idx["tempo"] - 1, vocab_size]
    for token in range(5):
        distributions[token - 1, start[token]:end[token]] = 1
    distributions[2, start[4]:end[4]] = 1
    length_start = start_idx["length"]
    length_end = start_idx["time"] - 1
    distributions[1, length_start:length_end] *= length_tensor
    distributions[4, start_idx["pitch"]:start_idx["dyn"] - 1] *= 10
    return distributions


def pick_distributions_by_prev_token(input_tokens, disc=DISCRETIZATION):
    """train.py:114-131."""
    _, start_idx = vocab_layout(disc)
    boundaries = [start_idx["dyn"] - 1, start_idx["length"] - 1, start_idx["time"] - 1, start_idx["tempo"] - 1]
    bins = torch.tensor(boundaries, device=input_tokens.device)
    buckets = torch.bucketize(input_tokens, bins, right=False)
    distributions = make_distributions(input_tokens.device, disc)
    buckets = buckets.long().to(distributions.device)
    return distributions[buckets]


def filtered_logit(input, output, disc=DISCRETIZATION):
    """train.py:133-138."""
    weights = pick_distributions_by_prev_token(input, disc)
    log_probs = F.log_softmax(output, dim=1)
    return -log_probs * weights


def loss_fn(src, trg, output, disc=DISCRETIZATION):
    """train.py:161-165."""
    vocab_size, _ = vocab_layout(disc)
    filtered_output = filtered_logit(src, output, disc)
    filtered_output = filtered_output.reshape(-1, vocab_size)
    return torch.nn.CrossEntropyLoss()(filtered_output, trg.reshape(-1))


def generate_greedy(model, context_len, token_ids, meta_ids, num_tokens, disc=DISCRETIZATION):
    """scripts/generate_midi_many.py:13-56 (argmax at :46), batch of 1 as in the script."""
    _, start_idx = vocab_layout(disc)
    model.eval()
    generated = token_ids.detach().cpu().numpy().tolist()[0]
    with torch.no_grad():
        for _ in range(num_tokens):
            logits = model(token_ids, meta_ids)
            filtered_logits = filtered_logit(token_ids, logits, disc)
            logits_last = filtered_logits[:, -1, :]
            if len(generated) > 0:
                recent = generated[-100:]
                counts = Counter(recent)
                for token, count in counts.items():
                    if start_idx["tempo"] <= token:
                        continue
                    elif start_idx["time"] <= token:
                        penalty = 1.1 * count if count >= 10 else 1
                    elif start_idx["length"] <= token:
                        penalty = min(1.015 ** count, 1.08)
                    elif start_idx["dyn"] <= token:
                        continue
                    else:
                        penalty = min(1.04 ** count, 1.25)
                    if count > 0:
                        logits_last[0, token] /= penalty
            next_token = logits_last.argmax(-1).unsqueeze(0)
            generated.append(next_token.item())
            token_ids = torch.cat([token_ids, next_token], dim=1)
            token_ids = token_ids[:, -context_len:]
    return generated


def choose_sampling(logits_last_row, cur_gen, u0, u1, start_idx):
    """One row of scripts/generate.py:33-85 — look-back window, k, penalties, top-k, one draw — with the two random
    numbers injected: `random.choice(seq)` is seq[int(u0 * len(seq))] and `torch.multinomial(p, 1)` is the inverse CDF
    of p at u1 (the library calls consume their generators differently; identical uniforms -> identical tokens is the
    parity definition for this path).  logits_last_row: 1-D tensor, modified in place as the script does."""
    val = 0
    j = 0
    for j, token in enumerate(reversed(cur_gen)):                      # :38-44
        if start_idx["time"] <= token < start_idx["tempo"]:
            val += token - start_idx["time"]
        if val >= 64 * 16:
            break
    recent = cur_gen[-j:]                                              # :45-46 (j == 0: the whole list)

    def choice(seq):
        return seq[min(int(u0 * len(seq)), len(seq) - 1)]

    k = 1
    if start_idx["tempo"] <= cur_gen[-1]:                              # :48-58
        k = choice([1, 1, 1, 2, 2])
    elif start_idx["time"] <= cur_gen[-1]:
        pass
    elif start_idx["length"] <= cur_gen[-1]:
        pass
    elif start_idx["dyn"] <= cur_gen[-1]:
        k = choice([1, 3])
    else:
        k = choice([1, 2])
    for token, count in Counter(recent).items():                       # :60-73
        if start_idx["tempo"] <= token:
            continue
        elif start_idx["time"] <= token:
            continue
        elif start_idx["length"] <= token:
            continue
        elif start_idx["dyn"] <= token:
            penalty = min(1.02 ** count, 1.2)
        else:
            penalty = min(1.01 ** count, 1.2)
        if count > 0:
            logits_last_row[token] /= penalty
    topk_probs, topk_indices = torch.topk(logits_last_row, k)          # :78-81
    topk_probs = topk_probs / topk_probs.sum()
    acc, pick = 0.0, k - 1
    acc = torch.zeros((), dtype=topk_probs.dtype)
    for jj in range(k):
        acc = acc + topk_probs[jj]
        if u1 < float(acc):
            pick = jj
            break
    return int(topk_indices[pick])


def generate_sampling(model, context_len, token_ids, meta_ids, num_tokens, uniforms, disc=DISCRETIZATION):
    """scripts/generate.py:14-95 with injected uniforms [num_tokens, B, 2] (see choose_sampling)."""
    _, start_idx = vocab_layout(disc)
    model.eval()
    gen = [row.tolist() for row in token_ids.cpu()]
    with torch.no_grad():
        for step in range(num_tokens):
            if token_ids.size(1) > context_len:
                token_ids = token_ids[:, -context_len:]
            logits = model(token_ids, meta_ids)
            logits_last = filtered_logit(token_ids, logits, disc)[:, -1, :]
            nxt = []
            for i in range(len(gen)):
                tok = choose_sampling(logits_last[i], gen[i], float(uniforms[step, i, 0]), float(uniforms[step, i, 1]), start_idx)
                gen[i].append(tok)
                nxt.append(tok)
            token_ids = torch.cat([token_ids, torch.tensor(nxt, device=token_ids.device).unsqueeze(1)], dim=1)
    return gen
