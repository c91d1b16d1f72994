#!/usr/bin/env python
"""Golden vectors for the CIDEr-D self-critical reward, from the REFERENCE's own code (build container only).

``get_self_critical_reward`` is taken as source text from /root/reference/Utils.py (the module itself cannot be imported:
matplotlib / skimage / pycocotools / nltk are missing) and executed against the reference's own scorer, imported from
/root/reference/cider/pyciderevalcap/ciderD.  The one shim: ``CiderD(df='<dataset>-train')`` would unpickle
``cider/data/<dataset>-train.p`` (not shipped); the shim hands the scorer the same two fields -- ``document_frequency`` and
``ref_len`` -- computed by the reference's PreProcess/CIDEr_idf_preproccess.py functions from the synthetic corpus.
Corpus and rollouts are regenerated from seeds by ``simpleimagecaptionzoo_b200.synth``; only tokens, scores and rewards
are stored.  Usage: python tests/golden/make_golden_cider.py
"""
import json
import os
import re
import sys
import types
from collections import defaultdict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
sys.path.insert(0, REF)

from simpleimagecaptionzoo_b200 import synth  # noqa: E402
from cider.pyciderevalcap.ciderD.ciderD import CiderD  # noqa: E402  (the reference's scorer)


def load_idf_preprocess():
    src = open(os.path.join(REF, "PreProcess", "CIDEr_idf_preproccess.py")).read()
    mod = types.ModuleType("ref_idf")
    exec(compile(src.split("def main(args):")[0], "CIDEr_idf_preproccess.py", "exec"), mod.__dict__)
    return mod


def load_reward_fn(df, ref_len):
    src = open(os.path.join(REF, "Utils.py")).read()
    m = re.search(r"def get_self_critical_reward\(.*?\n(?=#-+sample utils)", src, re.S)
    assert m, "get_self_critical_reward not found"

    def shim(df=None, **kw):  # CiderD(df='<dataset>-train') without the pickle
        c = CiderD(df="corpus")
        c.cider_scorer.df_mode = "synthetic-train"
        c.cider_scorer.ref_len = np.log(float(ref_len))
        c.cider_scorer.document_frequency = defaultdict(float, df_table)
        return c

    df_table = df
    ns = {"np": np, "torch": torch, "CiderD": shim}
    exec(compile(m.group(0), "Utils.py", "exec"), ns)
    return ns["get_self_critical_reward"]


class Vocab:
    def __init__(self, ix2word):
        self.ix2word = dict(enumerate(ix2word))


def run_case(name, n_images, vocab, n_batch, n_per_image, seed):
    ix2word, refs = synth.make_caption_corpus(n_images, vocab, seed=seed)
    idf = load_idf_preprocess()
    df = idf.compute_doc_freq(idf.create_crefs(refs))  # PreProcess/CIDEr_idf_preproccess.py:41-66
    ref_len = n_images                                   # :78
    img_index = list(np.random.Generator(np.random.PCG64(seed)).choice(n_images, size=n_batch, replace=False))
    gen, greedy = synth.make_rollouts(ix2word, refs, img_index, n_per_image, 20, seed)
    fn = load_reward_fn(df, ref_len)
    gts = {i: refs[i] for i in range(n_images)}
    # the reference scores one sample per image; several samples per image = the same call with the image repeated
    ids_rep = [int(i) for i in np.repeat(img_index, n_per_image)]
    greedy_rep = np.repeat(greedy, n_per_image, axis=0)
    rewards = fn(gen_result=torch.from_numpy(gen).long(), greedy_res=torch.from_numpy(greedy_rep).long(), ground_truth=gts,
                 img_ids=ids_rep, caption_vocab=Vocab(ix2word), dataset_name="synthetic")
    out = dict(gen=gen, greedy=greedy, img_index=np.asarray(img_index, np.int32), rewards=rewards.numpy().astype(np.float32),
               meta=np.array(json.dumps(dict(name=name, n_images=n_images, vocab=vocab, n_batch=n_batch, n_per_image=n_per_image,
                                             seed=seed, max_len=20, df_entries=len(df)))))
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    r = rewards.numpy()[:, 0]
    print(f"{name}: B={n_batch} n={n_per_image} df={len(df)} reward range {r.min():.3f}..{r.max():.3f} nonzero={np.mean(r != 0):.2f}")


if __name__ == "__main__":
    os.chdir(REF)
    run_case("cider_n1", 60, 64, 24, 1, 0)
    run_case("cider_n5", 80, 48, 16, 5, 1)
