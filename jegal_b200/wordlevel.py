"""Word-level pooling of the content branch — the host-side mirror of
JEGAL.get_word_level_embs (models/jegal.py:131-211), JEGAL.get_audio_word_level_embs (:213-252)
and JEGAL.pad_wordlevel_embs (:254-272), with the same names, arguments, return values and error
behaviour, computed by ONE launch of K5 (csrc/segmean.cu) per feature tensor instead of a Python
loop with one slice + mean + stack per word.

The reference walks the words on the host to build index ranges and does a tiny tensor op per
word; here the host walk only produces (begin, end) row ranges (pure integer work on the tokenizer
offsets / word boundaries, exactly the reference's rules, quirks included), and the arithmetic is
one kernel over the flattened [B * L, D] feature matrix.  No CPU fallback: the tensors must be on
an sm_100 device.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import ops
from .ops import JegalError

# tokenizer.cls_token_id, sep_token_id, pad_token_id of xlm-roberta (models/jegal.py:13,136); the reference
# reads them from a module-level tokenizer, the mirror takes them as an argument.
XLMR_SPECIAL_TOKEN_IDS = (0, 2, 1)


def _word_start_indices(offsets_b, ids_b, special: Sequence[int]) -> List[int]:
    """models/jegal.py:146-149: token i starts a word iff its offset starts at 0 and it is not special."""
    return [i for i, off in enumerate(offsets_b) if int(off[0]) == 0 and int(ids_b[i]) not in special]


def _audio_range(wb_b, idx: int, n_frames: int) -> Tuple[int, int]:
    """models/jegal.py:188-191 / :230-233: audio_emb[b, start:end+1] with Python slice semantics."""
    actual_start = int(wb_b[0][1])
    start, end = int(wb_b[idx][1]) - actual_start, int(wb_b[idx][2]) - actual_start
    lo, hi, _ = slice(start, end + 1).indices(n_frames)
    if hi <= lo:  # the reference then evaluates word_audio_embeddings[0] on an empty tensor
        raise IndexError("index 0 is out of bounds for dimension 0 with size 0")
    return lo, hi


def _pool(emb: torch.Tensor, ranges: List[Tuple[int, int]], counts: List[int]) -> List[torch.Tensor]:
    """One K5 launch over emb viewed as [B * L, D]; returns one [n_words, D] tensor per entry of counts."""
    if not emb.is_cuda:
        raise JegalError("word-level pooling runs on the GPU only (no CPU fallback)")
    if not ranges:
        return []
    x = emb.contiguous().view(-1, emb.shape[-1])
    r = torch.from_numpy(np.asarray(ranges, dtype=np.int32).T.copy()).to(emb.device, non_blocking=True)
    out = ops.segment_mean(x, r[0].contiguous(), r[1].contiguous())
    return list(torch.split(out, counts, dim=0))


def get_word_level_embs(text_emb, text, input_ids, offset_mapping, audio_emb=None, word_boundaries=None,
                        special_token_ids: Sequence[int] = XLMR_SPECIAL_TOKEN_IDS):
    """Same contract as JEGAL.get_word_level_embs (models/jegal.py:131-211): returns
    (word_text_emb, word_audio_emb, invalid_sample_idx); a word's text embedding is the mean of its
    sub-word tokens (the last word's range runs to the end of the padded sequence, :170-171), its audio
    embedding the mean of frames start..end inclusive relative to the clip's first word (:188-196)."""
    batch_size, seq_len = int(input_ids.shape[0]), int(input_ids.shape[1])
    ids = input_ids.detach().cpu().numpy() if isinstance(input_ids, torch.Tensor) else np.asarray(input_ids)
    offs = offset_mapping.detach().cpu().numpy() if isinstance(offset_mapping, torch.Tensor) else offset_mapping
    special = tuple(int(s) for s in special_token_ids if s is not None)
    n_frames = int(audio_emb.shape[1]) if audio_emb is not None else 0
    t_ranges, a_ranges, counts, invalid = [], [], [], []
    for b in range(batch_size):
        starts = _word_start_indices(offs[b], ids[b], special)
        tr, ar, valid = [], [], True
        for idx, _word in enumerate(text[b]):
            if idx >= len(starts):  # more words than word starts (:162-166)
                valid = False
                invalid.append(b)
                break
            hi = starts[idx + 1] if idx < len(starts) - 1 else seq_len
            tr.append((b * seq_len + starts[idx], b * seq_len + hi))
            if audio_emb is not None:
                lo_a, hi_a = _audio_range(word_boundaries[b], idx, n_frames)
                ar.append((b * n_frames + lo_a, b * n_frames + hi_a))
        if valid:
            if len(tr) <= 0:
                invalid.append(b)
            else:
                t_ranges += tr
                a_ranges += ar
                counts.append(len(tr))
    word_text_emb = _pool(text_emb, t_ranges, counts)
    word_audio_emb = _pool(audio_emb, a_ranges, counts) if audio_emb is not None else []
    return word_text_emb, word_audio_emb, invalid


def get_audio_word_level_embs(audio_emb, word_boundaries, invalid_sample_idx=None):
    """Same contract as JEGAL.get_audio_word_level_embs (models/jegal.py:213-252)."""
    batch_size, n_frames = int(audio_emb.shape[0]), int(audio_emb.shape[1])
    ranges, counts = [], []
    for b in range(batch_size):
        if invalid_sample_idx is not None and b in invalid_sample_idx:
            continue
        ar = [tuple(b * n_frames + v for v in _audio_range(word_boundaries[b], idx, n_frames))
              for idx in range(len(word_boundaries[b]))]
        if len(ar) > 0:
            ranges += ar
            counts.append(len(ar))
        elif invalid_sample_idx is not None:
            invalid_sample_idx.append(b)
        else:
            invalid_sample_idx = [b]
    return _pool(audio_emb, ranges, counts), invalid_sample_idx


def pad_wordlevel_embs(wordlevel_embs):
    """JEGAL.pad_wordlevel_embs (models/jegal.py:254-272): zero-pad to the longest clip and stack;
    the 'mask' is the list of valid lengths.  Plain tensor plumbing (one pad_sequence)."""
    lengths = [int(e.size(0)) for e in wordlevel_embs]
    padded = torch.nn.utils.rnn.pad_sequence(list(wordlevel_embs), batch_first=True, padding_value=0.0)
    return padded, lengths


def fuse_concat(word_audio_emb: List[torch.Tensor], word_text_emb: List[torch.Tensor]) -> List[torch.Tensor]:
    """torch.cat((audio, text), dim=-1) per clip (models/jegal.py:405-406, fusion_strategy='concat')."""
    return [torch.cat((a, t), dim=-1) for a, t in zip(word_audio_emb, word_text_emb)]
