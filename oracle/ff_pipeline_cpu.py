"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch fp32) of one whole FreeFine edit: DDIM inversion -> TCA/MMSA
sampling with local CFG and masked DDPM steps, built from the per-function restatements of oracle/ff_oracle.py.

Follows reference src/demo/model.py: FreeFine_generation :1012-1049, DDIM_inversion_func :1342-1364, invert
:817-925, Details_Preserving_regeneration :1640-1700, forward_sampling :476-622; attention dispatch of
src/utils/attention.py ca_forward :350-418 with the controller counter logic (:1051-1058, :1086-1090).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this file.
Pinned by tests/test_oracle_pipeline.py against tests/golden/pipeline.npz (latents the UNMODIFIED reference
produced for the same seeded inputs).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ff_oracle as O


class _State:
    def __init__(self):
        self.reset()
        self.num_att_layers = 0
        self.fg_retain = self.fg_ref = self.region = None

    def reset(self):
        self.cur_att_layer = 0
        self.use_tca = False
        self.use_style_align = False
        self.local_edit = False
        self.method = None
        self.cg = None
        self.layer_idx = list(range(16))

    def advance(self):
        self.cur_att_layer += 1
        if self.cur_att_layer == self.num_att_layers:
            self.cur_att_layer = 0


class OraclePipeline:
    def __init__(self, parts, noise_fn=None):
        """parts: freefine_b200.standin.build_standin(...) on CPU fp32; noise_fn(k, shape) -> k-th randn draw."""
        self.unet, self.vae = parts.unet, parts.vae
        self.tokenizer, self.text_encoder = parts.tokenizer, parts.text_encoder
        self.alphas = O.make_alphas_cumprod()
        self.st = _State()
        self.noise_fn = noise_fn
        self._k = 0
        n = 0
        for name, net in self.unet.named_children():
            place = "down" if "down" in name else ("up" if "up" in name else ("mid" if "mid" in name else None))
            if place is None:
                continue
            for m in net.modules():
                if m.__class__.__name__ == "Attention":
                    m.forward = self._make_forward(m, place)
                    n += 1
        self.st.num_att_layers = n

    # attention dispatch (reference ca_forward :388-404)
    def _make_forward(self, mod, place):
        st = self.st

        def forward(hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
            is_cross = encoder_hidden_states is not None
            ctx = encoder_hidden_states if is_cross else hidden_states
            q, k, v = mod.to_q(hidden_states), mod.to_k(ctx), mod.to_v(ctx)
            S = q.shape[1]
            if not is_cross and st.use_style_align:                  # every place is in style_align_scope (:388-389)
                src = O.process_mask_before_attention(st.fg_ref, S) if st.method == "sdsa" else None
                hs = O.style_align(q, k, v, mod.heads, mod.scale, src)
            elif not is_cross and st.use_tca and place == "up":
                if st.cur_att_layer // 2 not in st.layer_idx:
                    hs = O.plain_attention(q, k, v, mod.heads, mod.scale)
                else:
                    src = O.process_mask_before_attention(st.fg_ref, S)
                    tgt = O.process_mask_before_attention(st.fg_retain, S)
                    hs = O.tca(q, k, v, mod.heads, mod.scale, src, tgt, st.method, st.cg)
            elif st.local_edit and is_cross:
                region = O.process_mask_before_attention(st.region, S)
                hs = O.cross_local(q, k, v, mod.heads, mod.scale, region)
            else:
                hs = O.plain_attention(q, k, v, mod.heads, mod.scale)
            st.advance()
            return mod.to_out[0](hs)

        return forward

    def _text(self, prompts):
        ids = self.tokenizer(prompts, padding="max_length", max_length=77, return_tensors="pt").input_ids
        return self.text_encoder(ids)[0]

    def _noise(self, shape):
        if self.noise_fn is not None:
            t = self.noise_fn(self._k, shape)
        else:
            t = torch.randn(tuple(shape))
        self._k += 1
        return t

    @torch.no_grad()
    def invert(self, coarse, ori_img, num_step, start_step, max_steps=None):
        """invert (model.py:817-925) of [coarse, ori] with the empty prompt, guidance 1.0.  Returns latents list."""
        pre = lambda im: (torch.from_numpy(im).float() / 127.5 - 1).permute(2, 0, 1)[None]
        x = self.vae.encode(torch.cat([pre(coarse), pre(ori_img)]))["latent_dist"].mean * 0.18215
        emb = self._text(["", ""])
        ts = O.timesteps_for(num_step)
        out = [x]
        for i, t in enumerate(reversed(ts)):
            if i >= num_step - start_step or (max_steps is not None and i >= max_steps):
                continue
            eps = self.unet(x, t, emb)
            x, _ = O.inv_step(eps, int(t), x, self.alphas, num_step)
            out.append(x)
        self.st.reset()
        return out

    @torch.no_grad()
    def sample(self, inverted, prompt, tgt_mask, ori_mask, draw_mask, full_hw, num_step, start_step, end_step, gs, eta,
               method, use_auto_draw=False, cons_area=None, reduce_inp_artifacts=False, end_scale=0.5, max_steps=None):
        """Details_Preserving_regeneration + forward_sampling (model.py:1640-1700, :476-622).  Returns latents list."""
        st = self.st
        lat_h, lat_w = inverted[-1].shape[2:]
        fg, sh, ori, comp, lvar = O.prepare_various_mask(tgt_mask, ori_mask, draw_mask, full_hw[1], full_hw[0], lat_h, lat_w,
                                                         use_auto_draw, cons_area, reduce_inp_artifacts)
        st.reset()
        st.fg_retain, st.fg_ref, st.region = fg, ori, fg
        if method in ("ssa", "sdsa"):                                # model.py:514-520
            st.use_style_align, st.method = True, method
        else:
            st.use_tca, st.layer_idx = True, list(range(10, 16))
            st.method = "tca" if method == "tca" else "mmsa"
        st.local_edit = True
        refer = inverted[::-1]
        x = refer[0].clone()
        emb = torch.cat([self._text(["", ""]), self._text([prompt, ""])])
        ts = O.timesteps_for(num_step)
        out = [x]
        done = 0
        for i, t in enumerate(ts):
            if i < start_step:
                continue
            if max_steps is not None and done >= max_steps:
                break
            x = x.clone()
            x[1:] = refer[i - start_step + 1][1]
            if method == "tca":
                st.cg = O.linear_param(i, start_step, end_step, num_step, end_scale)
            elif method == "mmsa_es" and i >= end_step:
                st.use_tca = False
            eps4 = self.unet(torch.cat([x] * 2), t, emb)
            eu, ec = eps4.chunk(2)
            eps = O.cfg_local(eu, ec, gs, comp)
            noise = self._noise(eps.shape) if eta > 0 else None
            x, _ = O.ctrl_step(eps, int(t), x, lvar, eta, noise, self.alphas, num_step)
            out.append(x)
            done += 1
        st.reset()
        return out

    def decode(self, x):
        img = self.vae.decode(x / 0.18215)["sample"]
        img = (img / 2 + 0.5).clamp(0, 1)
        return (img.permute(0, 2, 3, 1).numpy() * 255).astype(np.uint8)
