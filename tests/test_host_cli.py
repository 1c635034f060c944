"""Host logic of the extractor / trainer command lines (no GPU): checkpoint key normalisation, sharded checkpoints,
the resume + rank-sharding work-list rule, the DataLoader-over-indices of main_release, and the data-parallel batch
split of the trainer on a world_size-2 gloo group."""
import json
import os
import random

import numpy as np
import pytest
import torch

from mertools_b200 import shard
from mertools_b200.extract import common


def test_legacy_layernorm_names_and_task_prefixes_are_normalised(tmp_path):
    """bert-base-chinese / bert-base-uncased still carry LayerNorm.gamma / .beta; from_pretrained renames them."""
    sd = {"bert.embeddings.LayerNorm.gamma": np.ones(4, np.float32), "bert.embeddings.LayerNorm.beta": np.zeros(4, np.float32),
          "bert.encoder.layer.0.output.LayerNorm.gamma": np.full(4, 2, np.float32),
          "bert.encoder.layer.0.output.dense.weight": np.eye(4, dtype=np.float32), "cls.predictions.bias": np.zeros(2, np.float32)}
    out = common.normalise_hf_keys(sd)
    assert set(out) == {"embeddings.LayerNorm.weight", "embeddings.LayerNorm.bias", "encoder.layer.0.output.LayerNorm.weight",
                        "encoder.layer.0.output.dense.weight", "cls.predictions.bias"}
    assert out["encoder.layer.0.output.LayerNorm.weight"][0] == 2
    torch.save({k: torch.from_numpy(v) for k, v in sd.items()}, tmp_path / "pytorch_model.bin")
    assert set(common.load_hf_state_dict(str(tmp_path))) == set(out)


def test_sharded_checkpoints_are_merged_through_their_index(tmp_path):
    from safetensors.numpy import save_file
    a = {"hubert.encoder.layers.0.attention.q_proj.weight": np.arange(6, dtype=np.float32).reshape(2, 3)}
    b = {"hubert.encoder.layers.1.attention.q_proj.weight": np.ones((2, 3), np.float32), "hubert.masked_spec_embed": np.zeros(3, np.float32)}
    save_file(a, str(tmp_path / "model-00001-of-00002.safetensors"))
    save_file(b, str(tmp_path / "model-00002-of-00002.safetensors"))
    wm = {k: "model-00001-of-00002.safetensors" for k in a}
    wm.update({k: "model-00002-of-00002.safetensors" for k in b})
    (tmp_path / "model.safetensors.index.json").write_text(json.dumps({"metadata": {}, "weight_map": wm}))
    sd = common.load_hf_state_dict(str(tmp_path))
    assert set(sd) == {"encoder.layers.0.attention.q_proj.weight", "encoder.layers.1.attention.q_proj.weight", "masked_spec_embed"}
    np.testing.assert_array_equal(sd["encoder.layers.0.attention.q_proj.weight"], a[next(iter(a))])
    with pytest.raises(AssertionError, match="no model.safetensors"):
        common.load_hf_state_dict(str(tmp_path / "nowhere"))


def test_do_normalize_follows_the_checkpoint(tmp_path):
    assert common.read_do_normalize(str(tmp_path)) is True          # no preprocessor_config.json: the HuBERT default
    (tmp_path / "preprocessor_config.json").write_text(json.dumps({"do_normalize": False, "sampling_rate": 16000}))
    assert common.read_do_normalize(str(tmp_path)) is False


def test_work_list_is_sharded_round_robin_and_resumes(tmp_path, monkeypatch):
    """rank r takes clips r::world of a rank-independent order; §5: skip what is already on disk."""
    items = [f"clip{i:03d}" for i in range(11)]
    random.Random(0).shuffle(items)
    out = lambda it: str(tmp_path / f"{it}.npy")  # noqa: E731
    for done in ("clip002", "clip007"):
        np.save(out(done), np.zeros(1))
    shares = []
    for rank in range(3):
        monkeypatch.setenv("RANK", str(rank))
        monkeypatch.setenv("WORLD_SIZE", "3")
        monkeypatch.setenv("LOCAL_RANK", str(rank))
        mine, r, w = shard.my_work(items, out)
        assert (r, w) == (rank, 3) and shard.device_index(5) == rank
        shares.append(mine)
    todo = sorted(set(items) - {"clip002", "clip007"})
    assert shares == [todo[0::3], todo[1::3], todo[2::3]]
    monkeypatch.setenv("MER_RESUME", "0")
    assert shard.my_work(items, out)[0] == sorted(items)[2::3]     # rank 2 of 3, nothing skipped
    monkeypatch.delenv("RANK"), monkeypatch.delenv("WORLD_SIZE"), monkeypatch.delenv("LOCAL_RANK")
    assert shard.device_index(5) == 5 and shard.my_work(items, out, resume=True)[0] == todo


def test_index_loaders_draw_what_the_reference_loaders_draw():
    """main_release.get_loaders = the loader classes of mer2023.py:31-79 over sample indices: with the same seed the
    batches are the ones a DataLoader over the reference's Data_Feat would yield (same sampler, same generator)."""
    from torch.utils.data import DataLoader, Dataset
    from torch.utils.data.sampler import SubsetRandomSampler

    from mertools_b200 import main_release as MR

    class Rows(Dataset):  # stands for Data_Feat: returns the sample's own index as its "name"
        def __len__(self):
            return 32

        def __getitem__(self, i):
            return {"name": i}

    random.seed(4)
    folds = MR.random_split_indexes(32, 5)
    torch.manual_seed(9)
    tr, ev, ts = MR.get_loaders(32, folds, [8, 8, 8], batch_size=8)
    got = [[b.tolist() for b in tr[0]], [b.tolist() for b in ev[0]], [b.tolist() for b in ts[1]], [b.tolist() for b in tr[0]]]
    torch.manual_seed(9)
    ref_tr = DataLoader(Rows(), batch_size=8, sampler=SubsetRandomSampler(folds[0][0]), collate_fn=lambda x: [r["name"] for r in x],
                        pin_memory=False)
    ref_ev = DataLoader(Rows(), batch_size=8, sampler=SubsetRandomSampler(folds[0][1]), collate_fn=lambda x: [r["name"] for r in x])
    ref_ts = DataLoader(Rows(), batch_size=8, shuffle=False, collate_fn=lambda x: [r["name"] for r in x])
    ref = [list(ref_tr), list(ref_ev), [b for b in ref_ts][:1], list(ref_tr)]
    assert got[0] == ref[0] and got[1] == ref[1] and got[3] == ref[3]
    assert got[2] == [list(range(8))]


def test_reference_init_is_torchs_own_constructor_sequence():
    """get_models(args) must start from what the reference's Attention(args) starts from: same values, same draws."""
    from mertools_b200.fusion import param_names, reference_init
    torch.manual_seed(5)
    sd = reference_init("utt", 768, 512, 1024, 128, 6, 1)
    after = torch.rand(1)
    assert list(sd) == param_names("utt") and sd["text_encoder.linear_1.weight"].shape == (128, 512)
    torch.manual_seed(5)
    nn = torch.nn
    ref = []
    for d in (768, 512, 1024, 384):
        ref += [nn.Linear(d, 128), nn.Linear(128, 128), nn.Linear(128, 128)]
    ref += [nn.Linear(128, 3), nn.Linear(128, 6), nn.Linear(128, 1)]
    assert torch.equal(after, torch.rand(1))                          # the generator advanced identically
    flat = [t for m in ref for t in (m.weight, m.bias)]
    assert all(torch.equal(a, b) for a, b in zip(sd.values(), flat))
    torch.manual_seed(5)
    sd = reference_init("frm_align", 768, 768, 768, 64, 6, 1)
    assert list(sd) == param_names("frm_align") and sd["audio_encoder.rnn.weight_hh_l0"].shape == (256, 64)


def _dp_worker(rank, world, port, out):
    import torch.distributed as dist

    from mertools_b200 import main_release as MR
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    x = torch.arange(26, dtype=torch.float32).view(26, 1) * (rank * 0 + 1)
    local = x[rank::world]
    full = MR._gather_strided(local, 26, rank, world)
    out[rank] = bool(torch.equal(full, x))
    dist.destroy_process_group()


def test_rank_strided_batches_reassemble_in_reference_order_gloo():
    """world_size-2 gloo: rows rank::2 of a 26-row batch (13 + 13) and of a 25-row one come back in batch order."""
    import socket

    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dp_worker, args=(2, port, out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}


def test_ragged_audio_launches_are_grouped_pipelined_and_restored_in_input_order():
    """AudioExtractor.extract_waves, ragged branch, with the encoder stubbed (host logic only): sorted clips are cut into
    launches of at most max_rows rows and max_samples padded samples, each row's tail is zeroed in the reused staging
    buffer, launch k + 1 is staged before launch k is read back, and the results come back in input order."""
    from mertools_b200.extract.audio import AudioExtractor

    calls = []

    class Enc:
        device = torch.device("cpu")
        hidden = 4

        def forward_ragged(self, rows, lens, normalize=True, want_frames=False):
            assert rows.shape == (len(lens), max(lens)) and normalize
            for r, n in enumerate(lens):
                assert float(rows[r, n:].abs().sum()) == 0.0          # stale samples of an earlier launch are gone
            calls.append(list(lens))
            utt = torch.stack([rows[r, :n].sum().repeat(4) for r, n in enumerate(lens)])      # a signature of the clip
            frames = [rows[r, :n].reshape(-1, 1)[:3].repeat(1, 4) for r, n in enumerate(lens)] if want_frames else None
            return utt, frames

    ext = object.__new__(AudioExtractor)
    ext.enc, ext.device, ext.max_rows, ext.ragged, ext.max_samples, ext.do_normalize = Enc(), torch.device("cpu"), 3, True, 40, True
    rng = np.random.default_rng(0)
    lens = [9, 3, 17, 5, 12, 20, 4, 8]
    waves = [rng.standard_normal(n) for n in lens]
    out = ext.extract_waves(waves, "UTTERANCE")
    assert calls == [[3, 4, 5], [8, 9, 12], [17, 20]]                 # sorted; <= 3 rows; rows * longest <= 40
    for w, o in zip(waves, out):
        assert o.shape == (4,) and abs(float(o[0]) - float(np.float32(w.astype(np.float32).sum()))) < 1e-4
    calls.clear()
    out = ext.extract_waves(waves, "FRAME")
    for w, o in zip(waves, out):
        assert o.shape == (3, 4) and np.allclose(o[:, 0], w[:3].astype(np.float32))
