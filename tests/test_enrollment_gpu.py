"""GPU parity for enrollment: the frozen trunk's 512-d output (`mmla_net_embed`) and the fit of the
transfer head (`mmla_head_fit`) against the oracle (torch-CPU trunk; torch-autograd restatement of Keras'
categorical_crossentropy + RMSprop from the same initial weights and the same sample order).

Tolerances: embeddings 2e-4 absolute (fp32 classifier path, tanh outputs in [-1, 1]); fitted weights after 40 epochs
(3 840 sequential RMSprop steps, fp32 on both sides with different summation orders) 2e-3 absolute on values of order
0.1, loss curve 1e-3 relative; predictions of the fitted heads agree on >= 99 % of the samples."""
import numpy as np
import pytest

from oracle import head_fit as ohf, nets as onets, psf, synth

pytestmark = pytest.mark.gpu


def _oracle_embed(x, w, spec):
    return onets.speaker_forward(x, w, spec, return_embedding=True)


def test_embed_matches_oracle_trunk(cuda):
    from mmla_audio_b200 import models, weights as W
    spec = W.SPEAKER_BASE
    w = W.synthetic_weights(spec, 99)
    pcm = synth.synth_clips(600, 24, 40960)
    x = np.concatenate([psf.input_feature_gen(pcm[i]) for i in range(24)]).astype(np.float32)
    ref = _oracle_embed(x, w, spec)
    assert ref.shape == (24, 512)
    for precision, atol in (("fp32", 2e-4), ("tf32", 5e-3)):
        m = models.Model(spec, w, precision=precision)
        got = m.embed_device(cuda.from_numpy(x).cuda()).cpu().numpy()
        assert np.abs(got - ref).max() <= atol, (precision, np.abs(got - ref).max())
        # the head applied to the embedding is the model's own prediction
        kk, bk = W.dense_keys(spec)
        z = got @ w[kk] + w[bk]
        p = np.exp(z - z.max(1, keepdims=True))
        p /= p.sum(1, keepdims=True)
        np.testing.assert_allclose(p, m.predict(x), atol=2e-5)


def test_head_fit_matches_oracle_fit(cuda):
    from mmla_audio_b200 import enrollment
    rng = np.random.default_rng(5)
    n, per = 6, 64
    centers = rng.standard_normal((n, 512)).astype(np.float32) * 0.5
    emb = np.tanh(np.concatenate([centers[j] + 0.6 * rng.standard_normal((per, 512)).astype(np.float32) for j in range(n)]))
    labels = np.repeat(np.arange(n), per)
    y = np.eye(n, dtype=np.float32)[labels]
    epochs = 40
    k0 = enrollment.glorot_uniform(512, n, seed=3)
    b0 = np.zeros(n, np.float32)
    order = enrollment.epoch_orders(len(emb), epochs, seed=3)
    # a short last mini-batch: 384 samples, batch 20 -> 19 full + one of 4
    k, b, loss = enrollment.fit_head(emb, y, epochs=epochs, batch_size=20, kernel0=k0, bias0=b0, order=order)
    rk, rb, rloss = ohf.fit_head(emb, y, k0, b0, order, batch_size=20)
    print("head fit: max |dW|", np.abs(k - rk).max(), "max |db|", np.abs(b - rb).max(), "loss", loss[0], "->", loss[-1])
    assert np.abs(k - rk).max() <= 2e-3 and np.abs(b - rb).max() <= 2e-3
    np.testing.assert_allclose(loss, rloss, rtol=1e-3)
    assert loss[-1] < loss[0]
    sig = lambda z: 1 / (1 + np.exp(-z))
    assert ((sig(emb @ k + b).argmax(1)) == (sig(emb @ rk + rb).argmax(1))).mean() >= 0.99
    assert (sig(emb @ k + b).argmax(1) == labels).mean() >= 0.9


def test_transfer_learning_drop_in_fits_and_saves(cuda, tmp_path):
    """config 5 end to end, small: make_feature_experiment over enrollment recordings -> transfer_learning (head fit on
    the frozen trunk) -> saved model reloads with the customized_dense head and labels the enrollment chunks."""
    from mmla_audio_b200 import enrollment, models, weights as W
    from mmla_audio_b200 import speaker_identification as si
    spec = W.SPEAKER_BASE
    base = models.Model(spec, W.synthetic_weights(spec, 99), precision="fp32")
    corpus = [("spk%d" % j, synth.synth_clips(800 + 40 * j, 1, 16000 * 12)[0]) for j in range(4)]
    x, y, speaker_id = si.make_feature_experiment(corpus)
    assert x.shape[1:] == (256, 39) and y.shape[1] == 4 and len(speaker_id) == 4
    acc, model = enrollment.transfer_learning(x, y, 1, 0, base, str(tmp_path / "experiment" / "model"), epochs=60)
    assert model.spec.n_classes == 4 and model.spec.head_activation == "sigmoid" and 0.0 <= acc <= 1.0
    back = models.load_model(str(tmp_path / "experiment" / "model"), precision="fp32")
    assert back.spec.n_classes == 4 and back.spec.head_activation == "sigmoid"
    np.testing.assert_allclose(back.predict(x[:5]), model.predict(x[:5]), atol=1e-6)
