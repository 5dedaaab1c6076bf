"""ORACLE (test infrastructure, never shipped or measured as the product).

Plain-PyTorch fp32 restatement of the eval-mode forward of the reference's
CNNRNNModel / CNNRNNModelLarge (reference models/cnn_rnn_model.py) driven
directly by a checkpoint ``state_dict`` (keys of SURVEY.md Appendix B).

It exists because /root/reference does not travel to the GPU box; it is pinned
against the real reference modules by oracle/make_golden.py, which runs the
untouched reference here and commits inputs/outputs under tests/golden/
(tests/test_oracle.py replays them).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _bn_eval(x, sd, p, eps=1e-5):
    # nn.BatchNorm2d in eval(): (x - mean) / sqrt(var + eps) * gamma + beta
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], training=False, eps=eps)


def _conv_bn(x, sd, conv, bn, padding):
    return _bn_eval(F.conv2d(x, sd[conv + ".weight"], sd[conv + ".bias"], padding=padding), sd, bn)


def _pool_f(x):
    # nn.MaxPool2d(kernel_size=(2, 1)): halves the frequency axis, floors odd sizes
    return F.max_pool2d(x, kernel_size=(2, 1))


def _bilstm(x, sd, prefix, hidden, layers, cache=None):
    """nn.LSTM(batch_first=True, bidirectional=True) in eval mode, fp32
    (reference cnn_rnn_model.py:45-52,212-228; inter-layer dropout is off).  ``cache``: a dict that keeps the built
    module between calls -- the reference builds its modules once (main.py:41-54), so a timed loop must too."""
    inp = x.shape[-1]
    key = (prefix, str(x.device))
    rnn = cache.get(key) if cache is not None else None
    if rnn is None:
        rnn = torch.nn.LSTM(inp, hidden, num_layers=layers, batch_first=True, bidirectional=True).to(x.device)
        own = rnn.state_dict()
        rnn.load_state_dict({k: sd[prefix + "." + k] for k in own})
        rnn.eval()
        if cache is not None:
            cache[key] = rnn
    out, _ = rnn(x.float())
    return out


def _residual_block(x, sd, p):
    # reference cnn_rnn_model.py:93-99
    identity = _conv_bn(x, sd, p + ".skip.0", p + ".skip.1", padding=0)
    out = F.relu(_conv_bn(x, sd, p + ".conv1", p + ".bn1", padding=(1, 1)))
    out = _conv_bn(out, sd, p + ".conv2", p + ".bn2", padding=(1, 1))
    return F.relu(out + identity)


def _features(x):
    # (B,C,F,T) -> (B,T,C*F), feature index c*F + f (reference :60-62, :292-294)
    B, C, Fq, T = x.shape
    return x.permute(0, 3, 1, 2).reshape(B, T, C * Fq)


@torch.no_grad()
def small_forward(sd, x, hidden_size, num_layers, lstm_cache=None):
    """CNNRNNModel.forward (reference cnn_rnn_model.py:57-74): (B,1,F,T)->(B,88,T)."""
    h = _pool_f(F.relu(_conv_bn(x, sd, "model.cnn.0", "model.cnn.1", (1, 1))))
    h = _pool_f(F.relu(_conv_bn(h, sd, "model.cnn.4", "model.cnn.5", (1, 1))))
    feats = _features(h)
    if feats.shape[1] == 0:
        return torch.zeros(x.shape[0], 88, 1)
    r = _bilstm(feats, sd, "model.rnn", hidden_size, num_layers, lstm_cache)
    return F.linear(r, sd["model.fc.weight"], sd["model.fc.bias"]).transpose(1, 2)


def _attention(x, sd, num_heads=8, clip=10.0):
    # reference cnn_rnn_model.py:118-139 (dropout is identity in eval)
    B, T, C = x.shape
    hd = C // num_heads
    qkv = F.linear(x, sd["model.attention.qkv.weight"], sd["model.attention.qkv.bias"])
    qkv = qkv.reshape(B, T, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    a = (q @ k.transpose(-2, -1)) * (hd ** -0.5)
    a = torch.clamp(a, min=-clip, max=clip)
    a = torch.softmax(a, dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, T, C)
    return F.linear(o, sd["model.attention.proj.weight"], sd["model.attention.proj.bias"])


@torch.no_grad()
def large_forward(sd, x, hidden_size, num_layers, use_attention=True,
                  use_onset_offset_heads=True, return_all_heads=False, return_internals=False, lstm_cache=None):
    """CNNRNNModelLarge.forward (reference cnn_rnn_model.py:262-349).  The checkpoint must be fp32 (as the reference's is)."""
    B = x.shape[0]
    internals = {}
    h = _pool_f(F.relu(_conv_bn(x, sd, "model.conv1.0", "model.conv1.1", (1, 1))))
    internals["conv1"] = h
    h = _pool_f(_residual_block(h, sd, "model.res_block1"))
    internals["res1"] = h
    h = _residual_block(h, sd, "model.res_block2")
    internals["res2"] = h
    h = _pool_f(F.relu(_conv_bn(h, sd, "model.freq_aware_conv.0", "model.freq_aware_conv.1", (3, 1))))
    internals["freq"] = h
    feats = _features(h)
    if feats.shape[1] == 0:
        z = torch.zeros(B, 88, 1)
        if use_onset_offset_heads and return_all_heads:
            return {"frame": z, "onset": z.clone(), "offset": z.clone()}
        return z
    main = _bilstm(feats, sd, "model.rnn_main", hidden_size, num_layers, lstm_cache)
    local = _bilstm(feats, sd, "model.rnn_local", hidden_size // 2, 1, lstm_cache)
    r = torch.cat([main, local], dim=-1)
    internals["rnn"] = r
    if use_attention:
        a = _attention(r, sd)
        r = F.layer_norm(r + a, (r.shape[-1],), sd["model.attention_norm.weight"],
                         sd["model.attention_norm.bias"], eps=1e-6)
        internals["attn_norm"] = r
    if use_onset_offset_heads:
        s = F.relu(F.linear(r, sd["model.shared_fc.weight"], sd["model.shared_fc.bias"]))
        heads = {n: F.linear(s, sd[f"model.{n}_head.weight"], sd[f"model.{n}_head.bias"]).transpose(1, 2)
                 for n in ("frame", "onset", "offset")}
        out = heads if return_all_heads else heads["frame"]
    else:
        out = F.linear(r, sd["model.fc.weight"], sd["model.fc.bias"]).transpose(1, 2)
    return (out, internals) if return_internals else out


@torch.no_grad()
def forward(sd, x, model_type, hidden_size, num_layers, use_attention=True,
            use_onset_offset_heads=True, return_all_heads=False):
    """TranscriptionModel.forward dispatch (reference transcription_model.py:91-108)."""
    mt = model_type.lower()
    if mt in ("cnn_rnn", "cnn+rnn"):
        return small_forward(sd, x, hidden_size, num_layers)
    if mt in ("cnn_rnn_large", "large"):
        return large_forward(sd, x, hidden_size, num_layers, use_attention, use_onset_offset_heads,
                             return_all_heads and use_onset_offset_heads)
    raise ValueError(f"Unknown model type: {model_type}")


@torch.no_grad()
def predict(sd, x, model_type, hidden_size, num_layers, threshold=0.5, **kw):
    """TranscriptionModel.predict for CNN-RNN types (reference :263-266)."""
    logits = forward(sd, x, model_type, hidden_size, num_layers, **kw)
    return (torch.sigmoid(logits) > threshold).float()
