"""Drop-in for ``svecalign.vecalign.vecalign`` (reference: svecalign/vecalign/vecalign.py): the
``align()`` entry point that ``seg_align/align.py:208-230`` calls per document pair, with the same
keyword arguments, defaults, side effects (``"%s:%s:%.6f"`` lines, vecalign.py:174-184) and soft
fixes; the numerics run on the GPU through ``dp_utils.vecalign``.
"""
import logging
import os
import math
import pickle
import sys
from pathlib import Path
from typing import List, Optional, Set, Tuple, Union

from .dp_utils import vecalign
from .embedding_utils import (make_doc_embedding, make_doc_embedding_device, read_in_embedding_rows,
                              read_in_embeddings)

logger = logging.getLogger("vecalign")


def make_alignment_types(max_alignment_size: int):
    """vecalign.py:154-162: all (n, m), n outer, n+m <= max; list order = DP tie-break priority."""
    return [(x, y) for x in range(1, max_alignment_size) for y in range(1, max_alignment_size)
            if x + y <= max_alignment_size]


def make_many_to_one_alignment_types(max_alignment_size: int):
    """vecalign.py:165-171."""
    return [(m, 1) for m in range(1, max_alignment_size + 1)]


def print_alignments(alignments, scores=None, src_lines=None, tgt_lines=None, ofile=sys.stdout):
    """vecalign.py:174-184 — the format parsed by file_utils.read_alignments (literal_eval)."""
    if scores is None:
        scores = [None] * len(alignments)
    for (x, y), s in zip(alignments, scores):
        print('%s:%s' % (x, y) if s is None else '%s:%s:%.6f' % (x, y, s), file=ofile)
        if src_lines is not None and tgt_lines is not None:
            print(' ' * 40, 'SRC: ', ' '.join(src_lines[i].replace('\n', ' ').strip() for i in x), file=ofile)
            print(' ' * 40, 'TGT: ', ' '.join(tgt_lines[i].replace('\n', ' ').strip() for i in y), file=ofile)


def load_ignore_index_file(path: Union[str, Path]) -> Set[Tuple[int, int]]:
    """vecalign.py:187-195."""
    res = set()
    with open(path) as fp:
        for line in fp:
            i, j = line.strip().split(" ")
            item = (int(i), int(j))
            assert item not in res, f"{path}, {item}"
            res.add(item)
    return res


def width_over2_for(src_max: int, tgt_max: int, search_buffer_size: int) -> int:
    """vecalign.py:243."""
    return math.ceil(max(src_max, tgt_max) / 2.0) + search_buffer_size


def read_alignments(fin):
    """utils/file_utils.py:80-98 (used for gold_alignment)."""
    from ast import literal_eval
    out = []
    with open(fin, 'rt', encoding="utf-8") as f:
        for line in f:
            fields = [x.strip() for x in line.split(':') if len(x.strip())]
            if len(fields) < 2:
                raise Exception('Got line "%s", which does not have at least two ":" separated fields' % line.strip())
            out.append((literal_eval(fields[0]), literal_eval(fields[1])))
    return out


def align(src: str, tgt: str, src_embed: List[str], src_stopes: bool, tgt_stopes: bool, tgt_embed: List[str],
          alignment_max_size: int, many_to_one: Optional[int], search_buffer_size: int,
          del_percentile_frac: float, max_size_full_dp: int, costs_sample_size: int, num_samps_for_norm: int,
          overlap_segments: bool, print_aligned_text: bool, src_fp16: bool = False, tgt_fp16: bool = False,
          src_ignore_indices: Optional[Union[str, Path]] = None,
          tgt_ignore_indices: Optional[Union[str, Path]] = None, verbose: bool = False,
          debug_save_stack: Optional[str] = None, gold_alignment: Optional[str] = None,
          print_results: bool = False, save_aligned_text_to_file: Optional[str] = None):
    """vecalign.py:198-293, same keywords.  Returns the stack (the reference returns None)."""
    if verbose:
        logger.setLevel(logging.DEBUG)
    if alignment_max_size < 2:
        logger.warning('Alignment_max_size < 2. Increasing to 2 so that 1-1 alignments will be considered')
        alignment_max_size = 2
    src_max = many_to_one if many_to_one is not None else alignment_max_size - 1
    tgt_max = 1 if many_to_one is not None else alignment_max_size - 1
    types = (make_many_to_one_alignment_types(many_to_one) if many_to_one is not None
             else make_alignment_types(alignment_max_size))
    width_over2 = width_over2_for(src_max, tgt_max, search_buffer_size)

    logger.info(f'Aligning src={src} to tgt={tgt}')
    src_lines = open(src, 'rt', encoding="utf-8").readlines()
    tgt_lines = open(tgt, 'rt', encoding="utf-8").readlines()
    src_ign = load_ignore_index_file(src_ignore_indices) if src_ignore_indices else None
    tgt_ign = load_ignore_index_file(tgt_ignore_indices) if tgt_ignore_indices else None
    if debug_save_stack or os.environ.get("SVX_HOST_GATHER"):
        # the reference's host path (embedding_utils.py:135-203): the pickled stack then holds host arrays
        src_map, src_rows = read_in_embeddings(src_embed[0], src_embed[1], src_stopes, src_fp16)
        tgt_map, tgt_rows = read_in_embeddings(tgt_embed[0], tgt_embed[1], tgt_stopes, tgt_fp16)
        vecs0 = make_doc_embedding(src_map, src_rows, src_lines, src_max, ignore_indices=src_ign, overlap_segments=overlap_segments)
        vecs1 = make_doc_embedding(tgt_map, tgt_rows, tgt_lines, tgt_max, ignore_indices=tgt_ign, overlap_segments=overlap_segments)
    else:
        # same tensors, bit for bit, gathered on the GPU from the rows in their on-disk dtype (svx_gather_doc_embedding)
        src_map, src_rows = read_in_embedding_rows(src_embed[0], src_embed[1], src_stopes, src_fp16)
        tgt_map, tgt_rows = read_in_embedding_rows(tgt_embed[0], tgt_embed[1], tgt_stopes, tgt_fp16)
        vecs0 = make_doc_embedding_device(src_map, src_rows, src_lines, src_max, ignore_indices=src_ign, overlap_segments=overlap_segments)
        vecs1 = make_doc_embedding_device(tgt_map, tgt_rows, tgt_lines, tgt_max, ignore_indices=tgt_ign, overlap_segments=overlap_segments)

    stack = vecalign(vecs0=vecs0, vecs1=vecs1, final_alignment_types=types,
                     del_percentile_frac=del_percentile_frac, width_over2=width_over2,
                     max_size_full_dp=max_size_full_dp, costs_sample_size=costs_sample_size,
                     num_samps_for_norm=num_samps_for_norm, debug=bool(debug_save_stack))

    if print_results:
        out = open(save_aligned_text_to_file, mode="w") if save_aligned_text_to_file else sys.stdout
        print_alignments(stack[0]['final_alignments'], scores=stack[0]['alignment_scores'],
                         src_lines=src_lines if print_aligned_text else None,
                         tgt_lines=tgt_lines if print_aligned_text else None, ofile=out)
        if save_aligned_text_to_file:
            out.close()
    if debug_save_stack:
        pickle.dump(stack, open(debug_save_stack, mode="wb"))
    if gold_alignment is not None:
        from .score import score_multiple, log_final_scores
        res = score_multiple(gold_list=[read_alignments(gold_alignment)], test_list=[stack[0]['final_alignments']])
        log_final_scores(res)
    return stack
