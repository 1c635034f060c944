"""Golden fixture for the VGGish branch: the reference's OWN graph definition and extractor,
unmodified -- audio/vggish/vggish_slim.py:define_vggish_slim + load_vggish_slim_checkpoint and
audio/extract_vggish_embedding.py:extract (wav -> vggish_input.wavfile_to_examples -> batched sess.run -> save rules)
-- executed over tests/golden/tf_slim_shim.py, a torch-backed stand-in for the TensorFlow / tf_slim calls those files
make (TensorFlow is not installed in the build container).  Pins the graph STRUCTURE and the extractor logic to the
reference's source; TensorFlow's own float arithmetic is out of reach.

Run once in the build container (needs /root/reference; NOT on the GPU box):
    python tests/golden/make_golden_vggish.py
Writes tests/golden/vggish_golden.npz.  Stubs: `config` (paths), `soundfile.read` via scipy, an empty `resampy`; the
"checkpoint" is synthetic.vggish_state_dict(seed=8) saved under the TF variable names."""
import os
import sys
import tempfile
import types

import numpy as np
from scipy.io import wavfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
REF = "/root/reference/MERBench/feature_extraction/audio"

from mertools_b200 import synthetic as S  # noqa: E402

CLIPS = {"clipA": (48000, 500), "clipB": (20011, 501), "clipC": (17000, 502)}  # name -> (samples, seed); C: one example


def main():
    import tf_slim_shim
    tf, _ = tf_slim_shim.install(96, 64)
    sys.modules.setdefault("resampy", types.ModuleType("resampy"))
    sf = types.ModuleType("soundfile")
    sf.read = lambda path, dtype=None: (wavfile.read(path)[1], wavfile.read(path)[0])
    sys.modules["soundfile"] = sf
    with tempfile.TemporaryDirectory() as tmp:
        cfg = types.ModuleType("config")
        cfg.PATH_TO_PRETRAINED_MODELS = os.path.join(tmp, "models")
        cfg.PATH_TO_RAW_AUDIO, cfg.PATH_TO_FEATURES = {}, {}
        sys.modules["config"] = cfg
        os.makedirs(os.path.join(cfg.PATH_TO_PRETRAINED_MODELS, "vggish"))
        sd = S.vggish_state_dict(seed=8)
        with open(os.path.join(cfg.PATH_TO_PRETRAINED_MODELS, "vggish", "vggish_model.ckpt"), "wb") as f:
            np.savez(f, **sd)
        wavs, out = [], {}
        for name, (n, seed) in CLIPS.items():
            w = S.synth_waves(1, n, seed=seed)[0].astype(np.int16)
            path = os.path.join(tmp, name + ".wav")
            wavfile.write(path, 16000, w)
            wavs.append(path)
        sys.path.insert(0, REF)
        import extract_vggish_embedding as ref_script   # the unmodified reference module
        from vggish import vggish_params, vggish_slim
        for level in ("UTTERANCE", "FRAME"):
            save = os.path.join(tmp, level)
            os.makedirs(save)
            ref_script.extract(wavs, save, level, batch_size=3)   # batch_size 3: the batching loop takes several turns
            for name in CLIPS:
                out[f"{name}_{level}"] = np.load(os.path.join(save, name + ".npy")).astype(np.float32)
        # the bare graph on random log-mel patches (what tests feed the CUDA network and the oracle restatement)
        x = np.random.default_rng(9).normal(-2.0, 2.0, (5, 96, 64)).astype(np.float32)
        with tf.Graph().as_default(), tf.Session() as sess:
            vggish_slim.define_vggish_slim(training=False)
            vggish_slim.load_vggish_slim_checkpoint(sess, os.path.join(cfg.PATH_TO_PRETRAINED_MODELS, "vggish",
                                                                         "vggish_model.ckpt"))
            names = sorted(v.name for v in tf.global_variables())
            [emb] = sess.run([sess.graph.get_tensor_by_name(vggish_params.OUTPUT_TENSOR_NAME)],
                             feed_dict={sess.graph.get_tensor_by_name(vggish_params.INPUT_TENSOR_NAME): x})
        out["patches"], out["patch_embeddings"] = x, emb.astype(np.float32)
    assert names == sorted(k + ":0" for k in sd), "the graph's variables are not the checkpoint names of the oracle"
    np.savez(os.path.join(HERE, "vggish_golden.npz"), clip_names=np.array(list(CLIPS)),
             clip_samples=np.array([v[0] for v in CLIPS.values()]), clip_seeds=np.array([v[1] for v in CLIPS.values()]),
             variable_names=np.array(names), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
