"""Host-side mirror of FreeFine's editing pipeline (reference src/demo/model.py, class FreeFinePipeline), with the
per-step arithmetic on the sm_100a kernels:

* every UNet attention -> ff_attn_masked_kv through the patched `Attention.forward` (freefine_b200/attention.py);
* local CFG + masked DDPM/DDIM update (model.py:605-617 + ctrl_step :134-198) -> ONE launch of ff_ddim_cfg_step;
* DDIM inversion update (inv_step :109-132) -> ff_ddim_inv_step.

Method names, argument meaning and error behaviour follow the reference (`inv_step`, `ctrl_step`, `invert`,
`forward_sampling`, `DDIM_inversion_func`, `Details_Preserving_regeneration`, `FreeFine_generation`,
`prepare_various_mask`, `prepare_tensor_mask`, `dilate_mask`, `linear_param`, `image2latent`, `latent2image`).
The UNet / VAE / text encoder bodies are whatever modules the caller supplies (diffusers in the reference's
environment, the random-init stand-ins of freefine_b200/standin.py here) -- library kernels on both sides.

Extension over the reference: E independent edits can share one stream batch (`[edit_0, ref_0, edit_1, ref_1, ...]`
latents, `[u_e,u_r,c_e,c_r]` x E UNet streams); the reference runs exactly one edit per call (model.py:594,
attention.py:1034).  Scalars derived from `alphas_cumprod` stay on the host in fp32 and are computed with the
reference's own expressions, so the fused step is bit-identical to the reference arithmetic on CPU.
"""
from __future__ import annotations

from copy import deepcopy

import numpy as np
import torch
import torch.nn.functional as F

from . import ops
from .attention import (Attention_Modulator, override_forward, register_attention_control,
                        register_attention_control_4bggen, register_attention_control_compose)


def randn_tensor(shape, generator=None, device=None, dtype=None):
    """diffusers.utils.torch_utils.randn_tensor as the reference uses it (model.py:186-188).  Tests replace this
    module attribute to feed both sides the same noise."""
    return torch.randn(tuple(shape), generator=generator, device=device, dtype=dtype)


def seed_everything(seed):
    torch.manual_seed(seed)
    np.random.seed(seed % (2 ** 32))
    return seed


class FreeFinePipeline:
    """Drop-in for the reference class of the same name (model.py:103).  Construct from parts:
    FreeFinePipeline(unet, vae, tokenizer, text_encoder, scheduler[, controller])."""

    def __init__(self, unet, vae, tokenizer, text_encoder, scheduler, controller=None, device=None):
        self.unet, self.vae, self.tokenizer = unet, vae, tokenizer
        self.text_encoder, self.scheduler = text_encoder, scheduler
        self.controller = controller
        self._device = torch.device(device) if device is not None else next(unet.parameters()).device
        self.method_type = None

    @classmethod
    def from_parts(cls, parts, controller=None, device=None):
        return cls(parts.unet, parts.vae, parts.tokenizer, parts.text_encoder, parts.scheduler, controller, device)

    @property
    def device(self):
        return self._device

    def modify_unet_forward(self):
        """reference model.py:105-106"""
        self.unet.forward = override_forward(self.unet)

    # ------------------------------------------------------------------------------------------------------------
    # schedule scalars: the reference's own fp32 expressions on the host table
    # ------------------------------------------------------------------------------------------------------------
    def _ratio(self):
        return self.scheduler.config.num_train_timesteps // self.scheduler.num_inference_steps

    def _get_variance(self, timestep, prev_timestep):
        """reference model.py:200-209 (note `>= 0` here vs `> 0` in ctrl_step)."""
        a_t = self.scheduler.alphas_cumprod[timestep]
        a_prev = self.scheduler.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.scheduler.final_alpha_cumprod
        return ((1 - a_prev) / (1 - a_t)) * (1 - a_t / a_prev)

    def _ctrl_coefs(self, timestep, eta):
        t = int(timestep)
        prev_t = t - self._ratio()
        a_t = self.scheduler.alphas_cumprod[t]
        a_prev = self.scheduler.alphas_cumprod[prev_t] if prev_t > 0 else self.scheduler.final_alpha_cumprod
        std = eta * self._get_variance(t, prev_t).to(torch.float32) ** 0.5
        stdt = torch.cat((std[None,], torch.zeros_like(std)[None,]))
        c_ddpm = ((1 - a_prev - stdt ** 2) ** 0.5)[0]
        return dict(sqrt_1m_at=float((1 - a_t) ** 0.5), sqrt_at=float(a_t ** 0.5), sqrt_ap=float(a_prev ** 0.5),
                    c_ddim=float((1 - a_prev) ** 0.5), c_ddpm=float(c_ddpm), sigma=float(std))

    def _inv_coefs(self, timestep):
        nxt = int(timestep)
        t = min(nxt - self._ratio(), 999)
        a_t = self.scheduler.alphas_cumprod[t] if t >= 0 else self.scheduler.final_alpha_cumprod
        a_next = self.scheduler.alphas_cumprod[nxt]
        return dict(sqrt_1m_at=float((1 - a_t) ** 0.5), sqrt_at=float(a_t ** 0.5), sqrt_an=float(a_next ** 0.5),
                    c_next=float((1 - a_next) ** 0.5))

    # ------------------------------------------------------------------------------------------------------------
    # steps
    # ------------------------------------------------------------------------------------------------------------
    def inv_step(self, model_output, timestep, x, eta=0., verbose=False):
        """reference model.py:109-132 -> ff_ddim_inv_step.  Returns (x_next, pred_x0)."""
        return ops.ddim_inv_step(model_output.float().contiguous(), x.float().contiguous(), want_pred_x0=True,
                                 **self._inv_coefs(timestep))

    def _var_mask(self, mask, n_edits, h, w):
        m = mask
        if m.dtype != torch.uint8:
            # the reference only runs with uint8 masks on this path (quirk Q2); other dtypes are rejected loudly
            raise TypeError(f"ctrl_step mask must be uint8 (reference quirk Q2), got {m.dtype}")
        m = m.reshape(-1, h, w)
        if m.shape[0] == 1 and n_edits > 1:
            m = m.expand(n_edits, h, w)
        if m.shape[0] != n_edits:
            raise ValueError(f"mask for {m.shape[0]} edits, latents for {n_edits}")
        return m.contiguous().to(self.device)

    def ctrl_step(self, model_output, timestep, x, mask, eta: float = 0.0, generator=None):
        """reference model.py:134-198 for the [edit, ref] stream pairs (model_output already guidance-combined)
        -> ff_ddim_step.  Returns (x_prev, pred_x0)."""
        if model_output.shape[0] % 2 or mask is None:
            raise NotImplementedError("ctrl_step mirrors the 2-stream [edit, ref] local-DDPM form of the reference")
        E = model_output.shape[0] // 2
        h, w = x.shape[-2:]
        noise = None
        if eta > 0:
            noise = randn_tensor(model_output.shape, generator=generator, device=model_output.device,
                                 dtype=torch.float32).contiguous()
        return ops.ddim_step(model_output.float().contiguous(), x.float().contiguous(), noise,
                             self._var_mask(mask, E, h, w), want_pred_x0=True, **self._ctrl_coefs(timestep, eta))

    def cfg_ctrl_step(self, noise_pred4, timestep, x, cfg_mask, var_mask, guidance_scale, eta=0.0, generator=None):
        """Fused model.py:605-617: local CFG on the 4 UNet streams of each edit + ctrl_step, one kernel launch.
        noise_pred4 [4E,C,h,w] = [u_e,u_r,c_e,c_r] x E; x [2E,C,h,w]; returns x_prev [2E,C,h,w]."""
        E = x.shape[0] // 2
        h, w = x.shape[-2:]
        noise = None
        if eta > 0:
            if isinstance(generator, (list, tuple)):
                # one generator per edit: every edit sees the reference's own stream (seed_everything(seed), then one
                # randn_tensor([2,C,h,w]) draw per step, model.py:1018,186-188) whatever the batch / shard layout
                if len(generator) != E:
                    raise ValueError(f"{len(generator)} generators for {E} edits")
                noise = torch.stack([randn_tensor((2,) + tuple(x.shape[1:]), generator=g, device=x.device, dtype=torch.float32)
                                     for g in generator]).reshape(x.shape).contiguous()
            else:
                noise = randn_tensor(x.shape, generator=generator, device=x.device, dtype=torch.float32).contiguous()
        cm = None if cfg_mask is None else self._var_mask(cfg_mask, E, h, w)
        return ops.ddim_cfg_step(noise_pred4.float().contiguous(), x.float().contiguous(), noise, cm,
                                 self._var_mask(var_mask, E, h, w), float(guidance_scale),
                                 **self._ctrl_coefs(timestep, eta))

    def linear_param(self, t, t1, t0, t2, end_scale=0.5):
        """reference model.py:438-455: piecewise-linear context guidance 1 -> end_scale -> 0."""
        if t < t1 or t > t2:
            raise ValueError(f"t must be in [{t1}, {t2}]")
        if t <= t0:
            return 1 + (end_scale - 1) / (t0 - t1) * (t - t1)
        return end_scale + (-end_scale / (t2 - t0)) * (t - t0)

    # ------------------------------------------------------------------------------------------------------------
    # encode / decode / text
    # ------------------------------------------------------------------------------------------------------------
    def preprocess_image(self, image, device):
        """reference model.py:1282-1288: uint8 HWC -> [-1,1] 1CHW"""
        image = torch.from_numpy(image).float() / 127.5 - 1
        return image.permute(2, 0, 1).unsqueeze(0).to(device)

    @torch.no_grad()
    def image2latent(self, image):
        """reference model.py:224-268 (tensor / ndarray branches)."""
        if isinstance(image, np.ndarray):
            if image.dtype == np.uint8:
                image = image.astype(np.float32) / 255.0 * 2 - 1
            image = torch.from_numpy(image)
        tensor = image.clone().detach()
        if tensor.ndim == 3:
            tensor = tensor.permute(2, 0, 1).unsqueeze(0)
        elif tensor.ndim != 4:
            raise ValueError(f"Invalid tensor shape: {tensor.shape}")
        tensor = tensor.to(device=self.device, dtype=next(self.vae.parameters()).dtype)
        return self.vae.encode(tensor)['latent_dist'].mean * 0.18215

    @torch.no_grad()
    def latent2image(self, latents, return_type='np'):
        """reference model.py:269-280"""
        latents = 1 / 0.18215 * latents.detach()
        image = self.vae.decode(latents.to(next(self.vae.parameters()).dtype))['sample']
        image = (image.float() / 2 + 0.5).clamp(0, 1)
        if return_type == 'np':
            image = (image.cpu().permute(0, 2, 3, 1).numpy()[0] * 255).astype(np.uint8)
        return image

    @torch.no_grad()
    def get_text_embeddings(self, prompt):
        ids = self.tokenizer(prompt, padding="max_length", max_length=77, return_tensors="pt").input_ids
        return self.text_encoder(ids.to(self.device))[0]

    def _unet(self, latents, t, text_embeddings):
        dt = self.unet.dtype if hasattr(self.unet, "dtype") else next(self.unet.parameters()).dtype
        if latents.is_cuda and not (torch.is_tensor(t) and t.is_cuda):
            # the timestep as a device tensor made by a fill kernel: an H2D copy from pageable memory (what
            # `timesteps.to(device)` inside the UNet would do) synchronises the stream first, i.e. drains the launch queue
            # once per UNet call
            t = torch.full((), int(t), dtype=torch.int64, device=latents.device)
        with ops.nvtx_range(f"ff.unet streams={latents.shape[0]}"):
            out = self.unet(latents.to(dt), t, encoder_hidden_states=text_embeddings.to(dt))
        return out.float()

    # ------------------------------------------------------------------------------------------------------------
    # DDIM inversion loop (reference model.py:817-925)
    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def invert(self, image, prompt, num_inference_steps=50, num_actual_inference_steps=None, guidance_scale=7.5,
               eta=0.0, return_intermediates=False, verbose=False, **kwds):
        batch_size = image.shape[0]
        if isinstance(prompt, list):
            if batch_size == 1:
                image = image.expand(len(prompt), -1, -1, -1)
        elif isinstance(prompt, str) and batch_size > 1:
            prompt = [prompt] * batch_size
        text_embeddings = self.get_text_embeddings(prompt)
        latents = self.image2latent(image).float()
        if guidance_scale > 1.:
            uncond = self.get_text_embeddings([""] * batch_size)
            text_embeddings = torch.cat([uncond, text_embeddings], dim=0)
            self.controller.use_cfg = True
        self.scheduler.set_timesteps(num_inference_steps)
        latents_list = [latents]
        for i, t in enumerate(reversed(self.scheduler.timesteps)):
            if num_actual_inference_steps is not None and i >= num_actual_inference_steps:
                continue
            with ops.nvtx_range(f"ff.invert step {i} t={int(t)}"):
                model_inputs = torch.cat([latents] * 2) if guidance_scale > 1. else latents
                noise_pred = self._unet(model_inputs, t, text_embeddings)
                if guidance_scale > 1.:
                    eu, ec = noise_pred.chunk(2, dim=0)
                    noise_pred = eu + guidance_scale * (ec - eu)
                latents, _ = self.inv_step(noise_pred, t, latents)
            latents_list.append(latents)
        if return_intermediates:
            return latents, latents_list
        return latents

    # ------------------------------------------------------------------------------------------------------------
    # sampling loop (reference model.py:476-622), E edits per call
    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_sampling(self, prompt, prompt_embeds=None, refer_latents=None, batch_size=1, end_step=None,
                         height=512, width=512, num_inference_steps=50, num_actual_inference_steps=None,
                         guidance_scale=7.5, latents=None, unconditioning=None, neg_prompt=None,
                         return_intermediates=False, eta=0.0, end_scale=0.5, local_var_reg=None,
                         completion_mask_cfg=None, local_edit_text=True, share_attn=True, method_type=None,
                         verbose=False, local_perturbation=True, generator=None, **kwds):
        """`generator`: None (global generator, as the reference), one torch.Generator, or a list with one generator per
        edit (batched edits: each edit draws its own [2,C,h,w] noise per step, independent of batch composition)."""
        self.method_type = method_type
        assert guidance_scale > 1.0, 'USING THIS MODULE CFG Must > 1.0'
        c = self.controller
        if share_attn:                                                    # reference :502-520
            if method_type == 'tca':
                c.use_tca, c.layer_idx, c.method = True, list(range(10, 16)), 'tca'
            elif method_type in ('mmsa', 'mmsa_es'):
                c.use_tca, c.layer_idx, c.method = True, list(range(10, 16)), 'mmsa'
            elif method_type == 'ssa':
                c.use_style_align, c.method = True, 'ssa'
            elif method_type == 'sdsa':
                c.use_style_align, c.method = True, 'sdsa'
        c.use_cfg = True
        c.local_edit = local_edit_text
        if prompt_embeds is None:
            if isinstance(prompt, str) and batch_size > 1:
                prompt = [prompt] * batch_size
            cond = self.get_text_embeddings(prompt)
        else:
            cond = prompt_embeds
        n2 = cond.shape[0]                                                # 2 streams [edit, ref] per edit
        if n2 % 2:
            raise ValueError("forward_sampling expects [edit_prompt, \"\"] pairs")
        E = n2 // 2
        if latents is None:
            latents = torch.randn((n2, self.unet.in_channels, height // 8, width // 8), device=self.device)
        uncond = self.get_text_embeddings([neg_prompt if neg_prompt else ""] * n2)
        # UNet stream order per edit: [u_e, u_r, c_e, c_r]  (reference: cat([uncond, cond]) for its single edit)
        text_embeddings = torch.cat([uncond.reshape(E, 2, *uncond.shape[1:]), cond.reshape(E, 2, *cond.shape[1:])],
                                    dim=1).reshape(4 * E, *cond.shape[1:])
        self.scheduler.set_timesteps(num_inference_steps)
        latents = latents.float().clone()
        latents_list = [latents]
        if num_actual_inference_steps is None:
            num_actual_inference_steps = num_inference_steps
        start_step = num_inference_steps - num_actual_inference_steps
        C, h, w = latents.shape[1:]
        var_mask = local_var_reg if local_perturbation else torch.ones_like(local_var_reg)
        for i, t in enumerate(self.scheduler.timesteps):
            if i < start_step:
                continue
            # ref stream <- the inversion latent one step cleaner than t (quirk Q5, reference :582-586)
            ref_latent = refer_latents[i - start_step + 1].reshape(E, 2, C, h, w)[:, 1]
            if latents.shape[0] == 2 * E:
                latents = latents.clone()
                latents.view(E, 2, C, h, w)[:, 1] = ref_latent
            else:
                latents = torch.stack([latents.reshape(E, C, h, w), ref_latent], 1).reshape(2 * E, C, h, w)
            if method_type == 'tca':
                c.context_guidance = self.linear_param(i, start_step, end_step, num_inference_steps, end_scale=end_scale)
            elif method_type == 'mmsa_es' and i >= end_step:
                c.use_tca = False
            lat = latents.view(E, 2, C, h, w)
            model_inputs = torch.cat([lat, lat], dim=1).reshape(4 * E, C, h, w)
            te = text_embeddings
            if unconditioning is not None and isinstance(unconditioning, list):
                te = text_embeddings.clone().reshape(E, 4, *cond.shape[1:])
                te[:, :2] = unconditioning[i].to(te.dtype)
                te = te.reshape(4 * E, *cond.shape[1:])
            c.log_mask = False
            with ops.nvtx_range(f"ff.sample step {i} t={int(t)}"):
                noise_pred = self._unet(model_inputs, t, te)
                latents = self.cfg_ctrl_step(noise_pred, t, latents, completion_mask_cfg if local_edit_text else None,
                                             var_mask, guidance_scale, eta=eta, generator=generator)
            latents_list.append(latents)
        image = self.latent2image(latents, return_type="pt")
        if return_intermediates:
            return image, latents_list
        return image, None

    # ------------------------------------------------------------------------------------------------------------
    # masks (reference model.py:927-934, :1432-1512, :1622-1639): integer work, bit-exact incl. uint8 wrap (Q1)
    # ------------------------------------------------------------------------------------------------------------
    def mask_reduce_dim(self, mask):
        return mask[:, :, 0] if mask.ndim == 3 else mask

    def dilate_mask(self, mask, dilate_factor=15):
        """cv2.dilate(mask.astype(uint8), ones(k,k)) (reference model.py:927-934): sliding max with the anchor at k//2,
        outside = 0.  Returns a numpy uint8 array like the reference.  On a CUDA device binary masks go through
        ff_dilate_mask; the CPU form (sliding max, also for non-binary masks) is the restatement the CPU tests pin."""
        k = int(dilate_factor)
        a = k // 2
        m = torch.from_numpy(np.ascontiguousarray(mask.astype(np.uint8))).to(self.device)
        squeeze = m.dim() == 2
        if m.is_cuda and m.shape[1] % 4 == 0 and int(m.max()) <= 1:
            y = ops.dilate_mask(m if squeeze else m.permute(2, 0, 1).contiguous(), k)
            return (y if squeeze else y.permute(1, 2, 0)).cpu().numpy()
        x = (m[None, None] if squeeze else m.permute(2, 0, 1)[None]).float()
        x = F.pad(x, (a, k - 1 - a, a, k - 1 - a), value=0.0)
        y = F.max_pool2d(x, kernel_size=k, stride=1)
        y = y[0, 0] if squeeze else y[0].permute(1, 2, 0)
        return y.to(torch.uint8).cpu().numpy()

    def prepare_tensor_mask(self, mask, sup_res_w, sup_res_h, binary=True):
        """reference model.py:1622-1639"""
        if mask.ndim == 3:
            mask = mask[:, :, 0]
        t = torch.tensor(mask, device=self.device)[None, None]
        t = F.interpolate(t, (sup_res_h, sup_res_w), mode="nearest").squeeze(0).squeeze(0)
        if binary:
            t[t > 0.0] = 1.0
        else:
            norm_ = t.max()
            t = t.float()
            t /= norm_
        return t

    def _masks_for_kernel(self, masks, sup_res_w, sup_res_h, init_code):
        """uint8 [1,H,W] device copies of the numpy masks when ff_mask_prep applies (CUDA, integer masks already at the
        working resolution, latent grid dividing it); None sends the caller down the general per-op path."""
        if self.device.type != 'cuda':
            return None
        lh, lw = init_code.shape[2], init_code.shape[3]
        if sup_res_h % lh or sup_res_w % lw or sup_res_w % 4:
            return None
        out = []
        for m in masks:
            if m is None:
                out.append(None)
                continue
            m = np.asarray(m)
            m = m[:, :, 0] if m.ndim == 3 else m
            if m.shape != (sup_res_h, sup_res_w) or m.dtype.kind not in 'ub':
                return None
            out.append(torch.from_numpy(np.ascontiguousarray(m.astype(np.uint8)))[None].to(self.device))
        return out

    def prepare_various_mask(self, shifted_mask, ori_mask, draw_mask, sup_res_w, sup_res_h, init_code, verbose=False,
                             use_auto_draw=False, cons_area=None, reduce_inp_artifacts=False):
        """reference model.py:1432-1512 -> (fg_mask, shifted, ori, completion [lat], local_var [lat]).  The tensors are
        uint8 and the algebra wraps exactly like the reference (cons - ori, 1 - x: quirk Q1)."""
        if use_auto_draw or reduce_inp_artifacts:
            assert cons_area is not None, 'for auto draw / auto artifact expansion use cons area '
        dev_masks = self._masks_for_kernel((shifted_mask, ori_mask, None if use_auto_draw else draw_mask,
                                            cons_area if (use_auto_draw or reduce_inp_artifacts) else None),
                                           sup_res_w, sup_res_h, init_code)
        if dev_masks is not None:               # CUDA: dilations + algebra + down-sampling in one launch (ff_mask_prep)
            r = ops.mask_prep(*dev_masks, (init_code.shape[2], init_code.shape[3]), use_auto_draw, reduce_inp_artifacts)
            return tuple(t[0] for t in r)
        P = lambda m: self.prepare_tensor_mask(m, sup_res_w, sup_res_h)
        if not use_auto_draw:
            if not reduce_inp_artifacts:
                sh, ori = P(shifted_mask), P(ori_mask)
                flex = P(draw_mask) * (1 - sh)
                fg = flex + sh
                fg[fg > 0] = 1.0
                comp, lvar = flex, flex
            else:
                assert cons_area is not None, 'for auto artifact expansion use cons area '
                dil = P(self.dilate_mask(ori_mask, 30))
                cons, sh, ori = P(cons_area), P(shifted_mask), P(ori_mask)
                flex = P(draw_mask) * (1 - sh)
                fg = flex + sh
                fg[fg > 0] = 1.0
                comp = flex
                lvar = (1 - cons) * (1 - sh) * dil + flex
                lvar[lvar > 0] = 1
        else:
            assert cons_area is not None, 'for auto draw better use cons area '
            if not reduce_inp_artifacts:
                dil_t = P(self.dilate_mask(shifted_mask, 15))
                sh, ori, cons = P(shifted_mask), P(ori_mask), P(cons_area)
                fg = sh
                cons = cons - ori
                comp = (1 - cons) * (1 - sh) * dil_t
                lvar = comp
            else:
                dil_t = P(self.dilate_mask(shifted_mask, 15))
                dil = P(self.dilate_mask(ori_mask, 30))
                sh, ori, cons = P(shifted_mask), P(ori_mask), P(cons_area)
                fg = sh
                cons = cons - ori
                comp = dil + dil_t
                comp[comp > 0] = 1
                comp *= (1 - cons) * (1 - sh)
                lvar = comp
        size = (init_code.shape[2], init_code.shape[3])
        comp = F.interpolate(comp[None, None], size, mode='nearest').squeeze(0).squeeze(0)
        lvar = F.interpolate(lvar[None, None], size, mode='nearest').squeeze(0).squeeze(0)
        return fg, sh, ori, comp, lvar

    # ------------------------------------------------------------------------------------------------------------
    # entry points (reference model.py:1342-1364, :1640-1700, :1012-1049)
    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def DDIM_inversion_func(self, img, mask, prompt, num_step, start_step=0, ref_img=None, verbose=False):
        source_image = self.preprocess_image(img, self.device)
        if ref_img is not None:
            source_image = torch.cat((source_image, self.preprocess_image(ref_img, self.device)))
        if mask.ndim == 3:
            mask = mask[:, :, 0]
        latents, latents_list = self.invert(source_image, prompt, guidance_scale=1.0, num_inference_steps=num_step,
                                            num_actual_inference_steps=num_step - start_step,
                                            return_intermediates=True, verbose=verbose)
        self.controller.reset()
        return np.asarray(mask, dtype=np.float32), latents_list

    def Details_Preserving_regeneration(self, source_image, inverted_latents, edit_prompt, shifted_mask, ori_mask,
                                        draw_mask, num_steps=100, start_step=30, end_step=10, eta=1,
                                        guidance_scale=7.5, share_attn=True, method_type='tca', verbose=False,
                                        local_text_edit=True, local_perturbation=True, return_intermediates=False,
                                        use_auto_draw=False, cons_area=None, use_share_attention=False,
                                        reduce_inp_artifacts=False, end_scale=0.5):
        start_latents = inverted_latents[-1]
        init_code_orig = deepcopy(start_latents)
        full_h, full_w = source_image.shape[:2]
        fg_retain, fg_retain_st2, fg_ref, comp_cfg, local_var = self.prepare_various_mask(
            shifted_mask, ori_mask, draw_mask, full_h, full_w, init_code_orig, verbose=verbose,
            use_auto_draw=use_auto_draw, cons_area=cons_area, reduce_inp_artifacts=reduce_inp_artifacts)
        c = self.controller
        c.fg_retain_mask = fg_retain.to(self.device)
        c.fg_retain_mask_st2 = fg_retain_st2.to(self.device)
        c.fg_ref_mask = fg_ref.to(self.device)
        c.local_edit_region = fg_retain.to(self.device)
        c.reset()
        c.log_mask = False
        gen_images, intermediates = self.forward_sampling(
            prompt=[edit_prompt, ""], refer_latents=inverted_latents[::-1], end_step=end_step, batch_size=2,
            latents=init_code_orig, guidance_scale=guidance_scale, num_inference_steps=num_steps,
            num_actual_inference_steps=num_steps - start_step, eta=eta, completion_mask_cfg=comp_cfg,
            local_var_reg=local_var, share_attn=share_attn, method_type=method_type, verbose=verbose,
            blending=local_text_edit, local_perturbation=local_perturbation, return_intermediates=return_intermediates,
            use_share_attention=use_share_attention, end_scale=end_scale)
        edit_gen_image, ref_gen_image = gen_images
        c.reset()
        to_u8 = lambda im: (im.permute(1, 2, 0).detach().cpu().numpy() * 255).astype(np.uint8)
        return to_u8(edit_gen_image), to_u8(ref_gen_image), intermediates

    def FreeFine_generation(self, ori_img, ori_mask, coarse_input, target_mask, guidance_text, guidance_scale, eta,
                            end_step=10, num_step=50, start_step=25, share_attn=True, method_type='tca',
                            local_text_edit=True, local_perturbation=True, verbose=True, return_ori=False, seed=42,
                            draw_mask=None, return_intermediates=False, use_auto_draw=False, cons_area=None,
                            reduce_inp_artifacts=False, end_scale=0.5):
        """reference model.py:1012-1049.  With return_intermediates the list of latents is kept in
        `self.last_intermediates` (the reference writes a GIF instead: debug-only, out of scope)."""
        methods = ['tca', 'ssa', 'sdsa', 'mmsa', 'mmsa_es']
        assert method_type in methods, f"check method type f{method_type}, which is not in {methods}"
        seed_everything(seed)
        ori_mask = self.mask_reduce_dim(ori_mask)
        target_mask = self.mask_reduce_dim(target_mask)
        if draw_mask is not None:
            draw_mask = self.mask_reduce_dim(draw_mask)
        shifted_mask, inverted_latent = self.DDIM_inversion_func(img=coarse_input, mask=target_mask, prompt="",
                                                                 num_step=num_step, start_step=start_step,
                                                                 ref_img=ori_img, verbose=verbose)
        edit_img, ref_img, intermediates = self.Details_Preserving_regeneration(
            coarse_input, inverted_latent, guidance_text, target_mask, ori_mask, draw_mask, num_steps=num_step,
            start_step=start_step, end_step=end_step, guidance_scale=guidance_scale, eta=eta, share_attn=share_attn,
            method_type=method_type, verbose=verbose, local_text_edit=local_text_edit,
            local_perturbation=local_perturbation, return_intermediates=return_intermediates, cons_area=cons_area,
            use_auto_draw=use_auto_draw, end_scale=end_scale, reduce_inp_artifacts=reduce_inp_artifacts)
        self.last_intermediates = intermediates
        if not return_ori:
            return edit_img
        return edit_img, ref_img


    # ------------------------------------------------------------------------------------------------------------
    # background generation / object removal (reference model.py:656-812, :1088-1118, :1611-1620, :1751-1804);
    # the attention side is register_attention_control_4bggen + Temporal_contextal_attention_bg
    # ------------------------------------------------------------------------------------------------------------
    def prepare_mask_bggen(self, mask, sup_res_w, sup_res_h, init_code):
        """reference model.py:1611-1620"""
        mask_tensor = self.prepare_tensor_mask(mask, sup_res_w, sup_res_h)
        lvar = F.interpolate(mask_tensor[None, None], (init_code.shape[2], init_code.shape[3]), mode='nearest')
        return mask_tensor, lvar.squeeze(0).squeeze(0)

    @torch.no_grad()
    def forward_sampling_background_gen(self, prompt, batch_size=1, end_step=None, height=512, width=512,
                                        num_inference_steps=50, num_actual_inference_steps=None, guidance_scale=7.5,
                                        latents=None, refer_latents=None, unconditioning=None, neg_prompt=None,
                                        return_intermediates=False, eta=0.0, local_var_reg=None, local_cfg_reg=None,
                                        local_text_edit=True, share_attn=True, method_type='tca', verbose=False,
                                        local_perturbation=True, end_scale=0.5, latent_blended=True,
                                        blend_range=(0, 40), **kwds):
        """reference model.py:656-812.  The inversion ran on the source image alone (1 stream per edit); each step the
        ref stream is refer_latents[i-start_step] (no +1 offset here: quirk Q5) and only the edit stream is kept."""
        assert guidance_scale > 1.0, 'USING THIS MODULE CFG Must > 1.0'
        self.method_type = method_type
        c = self.controller
        if share_attn:
            if method_type == 'tca':
                c.use_tca, c.layer_idx, c.method = True, list(range(10, 16)), 'tca'
            elif method_type in ('mmsa', 'mmsa_es'):
                c.use_tca, c.layer_idx, c.method = True, list(range(10, 16)), 'mmsa'
            elif method_type == 'ssa':
                c.use_style_align, c.method = True, 'ssa'
            elif method_type == 'sdsa':
                c.use_style_align, c.method = True, 'sdsa'
        c.use_cfg = True
        c.local_edit = local_text_edit
        if isinstance(prompt, str) and batch_size > 1:
            prompt = [prompt] * batch_size
        cond = self.get_text_embeddings(prompt)
        n2 = cond.shape[0]
        E = n2 // 2
        uncond = self.get_text_embeddings([neg_prompt if neg_prompt else ""] * n2)
        text_embeddings = torch.cat([uncond.reshape(E, 2, *uncond.shape[1:]), cond.reshape(E, 2, *cond.shape[1:])],
                                    dim=1).reshape(4 * E, *cond.shape[1:])
        self.scheduler.set_timesteps(num_inference_steps)
        if num_actual_inference_steps is None:
            num_actual_inference_steps = num_inference_steps
        start_step = num_inference_steps - num_actual_inference_steps
        C, h, w = latents.shape[1:]
        edit = latents.float().reshape(-1, C, h, w)[:E] if latents.shape[0] == E else latents.float().reshape(E, -1, C, h, w)[:, 0]
        latents_list = [latents]
        var_mask = local_var_reg if local_perturbation else torch.ones_like(local_var_reg)
        for i, t in enumerate(self.scheduler.timesteps):
            if i < start_step:
                continue
            ref = refer_latents[i - start_step].float().reshape(E, C, h, w)
            lat = torch.stack([edit, ref], dim=1)                                   # [E,2,C,h,w]
            if method_type == 'tca':
                c.context_guidance = self.linear_param(i, start_step, end_step, num_inference_steps, end_scale=end_scale)
            elif method_type == 'mmsa_es' and i >= end_step:
                c.use_tca = False
            model_inputs = torch.cat([lat, lat], dim=1).reshape(4 * E, C, h, w)
            c.log_mask = False
            noise_pred = self._unet(model_inputs, t, text_embeddings)
            out = self.cfg_ctrl_step(noise_pred, t, lat.reshape(2 * E, C, h, w), local_cfg_reg if local_text_edit else None,
                                     var_mask, guidance_scale, eta=eta)
            edit = out.reshape(E, 2, C, h, w)[:, 0]
            latents_list.append(edit if E > 1 else edit[0])
            last = out
        image = self.latent2image(last, return_type="pt")
        if return_intermediates:
            return image, latents_list
        return image, None

    def Details_Preserving_regeneration_background(self, ori_img, inverted_latents, edit_prompt, ori_mask, num_steps=100,
                                                   start_step=30, end_step=10, guidance_scale=3.5, eta=1, verbose=False,
                                                   local_text_edit=True, local_perturbation=True, end_scale=0.5,
                                                   return_intermediates=False, share_attn=True, method_type='tca',
                                                   latent_blended=True, blend_range=(0, 40)):
        """reference model.py:1751-1804"""
        init_code_orig = deepcopy(inverted_latents[-1])
        full_h, full_w = ori_img.shape[:2]
        mask_tensor, local_var_reg = self.prepare_mask_bggen(ori_mask, full_h, full_w, init_code_orig)
        c = self.controller
        c.fg_retain_mask = mask_tensor.to(self.device)
        c.local_edit_region = mask_tensor.to(self.device)
        c.reset()
        gen_images, intermediates = self.forward_sampling_background_gen(
            prompt=[edit_prompt, ""], end_step=end_step, batch_size=2, refer_latents=inverted_latents[::-1],
            latents=init_code_orig, guidance_scale=guidance_scale, num_inference_steps=num_steps,
            num_actual_inference_steps=num_steps - start_step, eta=eta, local_cfg_reg=local_var_reg,
            local_var_reg=local_var_reg, share_attn=share_attn, method_type=method_type, verbose=verbose,
            local_text_edit=local_text_edit, local_perturbation=local_perturbation,
            return_intermediates=return_intermediates, end_scale=end_scale, latent_blended=latent_blended,
            blend_range=blend_range)
        c.reset()
        edit = (gen_images[0].permute(1, 2, 0).detach().cpu().numpy() * 255).astype(np.uint8)
        return edit, intermediates

    def FreeFine_background_generation(self, ori_img, ori_mask, guidance_text, guidance_scale, eta, end_step=10,
                                       num_step=50, start_step=25, share_attn=True, method_type='tca',
                                       local_text_edit=True, local_perturbation=True, verbose=True, seed=42,
                                       return_intermediates=False, end_scale=0.5, latent_blended=False,
                                       blend_range=(0, 40)):
        """reference model.py:1088-1118 (needs register_attention_control_4bggen).  Extra kwargs the published drivers
        pass (`use_auto_draw`, `reduce_inp_artifacts`) are rejected by the reference signature too."""
        seed_everything(seed)
        ori_mask = self.mask_reduce_dim(ori_mask)
        _, inverted = self.DDIM_inversion_func(img=ori_img, mask=ori_mask, prompt="", num_step=num_step,
                                               start_step=start_step, ref_img=None, verbose=verbose)
        edit, intermediates = self.Details_Preserving_regeneration_background(
            ori_img, inverted, guidance_text, ori_mask, num_steps=num_step, start_step=start_step, end_step=end_step,
            guidance_scale=guidance_scale, eta=eta, share_attn=share_attn, method_type=method_type, verbose=verbose,
            end_scale=end_scale, local_text_edit=local_text_edit, local_perturbation=local_perturbation,
            return_intermediates=return_intermediates, latent_blended=latent_blended, blend_range=blend_range)
        self.last_intermediates = intermediates
        return edit

    # ------------------------------------------------------------------------------------------------------------
    # cross-image composition / appearance transfer (reference model.py:301-435, :1051-1086, :1367-1388, :1515-1609,
    # :1701-1750); attention side: register_attention_control_compose + Temporal_contextal_attention_compose
    # ------------------------------------------------------------------------------------------------------------
    def prepare_composition_masks(self, ori_mask_lists, tgt_mask_lists, sup_res_w, sup_res_h, init_code,
                                  dil_completion=False, dil_factor=15, draw_mask=None, appearance_transfer=False):
        """reference model.py:1515-1609 -> (tgt_masks [n_tgt+1,H,W], ori_masks [N,H,W], local_perturbation [lat],
        completion_mask_cfg [lat]); uint8 algebra as the reference."""
        P = lambda m: self.prepare_tensor_mask(m, sup_res_w, sup_res_h)
        lat = (init_code.shape[2], init_code.shape[3])
        down = lambda t: F.interpolate(t[None, None], lat, mode='nearest').squeeze(0).squeeze(0)
        ori = [P(m) for m in ori_mask_lists]
        tgt = []
        lp, fg = torch.zeros_like(ori[0]), torch.zeros_like(ori[0])
        if appearance_transfer:
            for shifted in tgt_mask_lists:
                d = P(self.dilate_mask(shifted, dil_factor))
                tgt.append(d)
                lp += d
            lp[lp > 0] = 1
            tgt.append(1 - lp)
            lp = down(lp)
            return torch.stack(tgt), torch.stack(ori), lp, deepcopy(lp)
        if draw_mask is None:
            for shifted in tgt_mask_lists:
                d = P(self.dilate_mask(shifted, dil_factor))
                sh = P(shifted)
                tgt.append(d if dil_completion else sh)
                fg += sh
                lp += d
            fg[fg > 0] = 1
            lp[lp > 0] = 1
            tgt.append(1 - fg if dil_completion else 1 - lp)
            lp = down(lp * (1 - fg))
            return torch.stack(tgt), torch.stack(ori), lp, (deepcopy(lp) if dil_completion else torch.zeros_like(lp))
        for i, shifted in enumerate(tgt_mask_lists):
            sh = P(shifted)
            d = P(draw_mask[i]) + sh
            d[d > 0] = 1
            tgt.append(d)
            fg += sh
            lp += d
        fg[fg > 0] = 1
        lp[lp > 0] = 1
        tgt.append(1 - lp)
        lp = down(lp * (1 - fg))
        return torch.stack(tgt), torch.stack(ori), lp, lp

    @torch.no_grad()
    def DDIM_inversion_func_compose(self, img, compose_imgs, prompt, num_step, start_step=0, verbose=False):
        """reference model.py:1367-1388: invert [coarse, ref_1..ref_N] together."""
        source = torch.cat([self.preprocess_image(im, self.device) for im in [img] + list(compose_imgs)])
        _, latents_list = self.invert(source, prompt, guidance_scale=1.0, num_inference_steps=num_step,
                                      num_actual_inference_steps=num_step - start_step, return_intermediates=True,
                                      verbose=verbose)
        self.controller.reset()
        return latents_list

    @torch.no_grad()
    def forward_sampling_compose(self, prompt, prompt_embeds=None, refer_latents=None, batch_size=1, end_step=None,
                                 height=512, width=512, num_inference_steps=50, num_actual_inference_steps=None,
                                 guidance_scale=7.5, latents=None, unconditioning=None, neg_prompt=None,
                                 return_intermediates=False, eta=0.0, end_scale=0.5, local_var_reg=None,
                                 local_edit_text=True, cfg_masks_tensor=None, share_attn=True, method_type=None,
                                 verbose=False, local_perturbation=True, **kwds):
        """reference model.py:301-435: UNet streams [e, r_1..r_N, c_e]; text batch = (N+1) unconditional + the regional
        prompts + ""; guidance from streams 0 and N+1; only the edit stream is stepped (ff_ddim_cfg_step_compose)."""
        self.method_type = method_type
        assert guidance_scale > 1.0, 'USING THIS MODULE CFG Must > 1.0'
        c = self.controller
        if share_attn:
            if method_type == 'tca':
                c.use_tca, c.layer_idx, c.method = True, list(range(10, 16)), 'tca'
            elif method_type in ('mmsa', 'mmsa_es'):
                c.use_tca, c.layer_idx, c.method = True, list(range(10, 16)), 'mmsa'
            elif method_type == 'ssa':
                c.use_style_align, c.method = True, 'ssa'
            elif method_type == 'sdsa':
                c.use_style_align, c.method = True, 'sdsa'
        c.use_cfg = True
        c.local_edit = local_edit_text
        prompt = list(prompt) + [""]
        c.prompt_length = len(prompt)
        cond = self.get_text_embeddings(prompt)
        n_lat = latents.shape[0]                                      # N+1 = [edit, ref_1..ref_N]
        uncond = self.get_text_embeddings([neg_prompt if neg_prompt else ""] * n_lat)
        text_embeddings = torch.cat([uncond, cond], dim=0)
        self.scheduler.set_timesteps(num_inference_steps)
        if num_actual_inference_steps is None:
            num_actual_inference_steps = num_inference_steps
        start_step = num_inference_steps - num_actual_inference_steps
        C, h, w = latents.shape[1:]
        edit = latents.float()[:1]
        latents_list = [latents]
        var_mask = local_var_reg if local_perturbation else torch.ones_like(local_var_reg)
        for i, t in enumerate(self.scheduler.timesteps):
            if i < start_step:
                continue
            refs = refer_latents[i - start_step + 1][1:].float()      # one step cleaner than t (quirk Q5)
            if method_type == 'tca':
                c.context_guidance = self.linear_param(i, start_step, end_step, num_inference_steps, end_scale=end_scale)
            elif method_type == 'mmsa_es' and i >= end_step:
                c.use_tca = False
            model_inputs = torch.cat([edit, refs, edit])              # [e, r_1..r_N, c_e]
            c.log_mask = False
            noise_pred = self._unet(model_inputs, t, text_embeddings)
            noise = None
            if eta > 0:
                noise = randn_tensor(edit.shape, device=edit.device, dtype=torch.float32).contiguous()
            edit = ops.ddim_cfg_step_compose(noise_pred.contiguous(), model_inputs.shape[0], edit.contiguous(), noise,
                                             self._var_mask(cfg_masks_tensor, 1, h, w) if local_edit_text else None,
                                             self._var_mask(var_mask, 1, h, w), float(guidance_scale),
                                             **self._ctrl_coefs(t, eta))
            latents_list.append(edit[0])
        image = self.latent2image(edit, return_type="pt")[0]
        if return_intermediates:
            return image, latents_list
        return image, None

    @torch.no_grad()
    def Details_Preserving_regeneration_compose(self, source_image, inverted_latents, edit_prompt_list, ori_mask_lists,
                                                tgt_mask_lists, draw_mask, num_steps=100, start_step=30, end_step=10,
                                                eta=1, guidance_scale=7.5, dil_completion=False,
                                                appearance_transfer=False, share_attn=True, method_type='tca',
                                                verbose=False, local_text_edit=True, local_perturbation=True,
                                                return_intermediates=False, use_share_attention=False, dil_factor=15,
                                                end_scale=0.5):
        """reference model.py:1701-1750.  The attention plans hold at most FF_MAX_PASS passes per (stream, head): N source
        images need N + 1 of them ('tca' self-attention and the regional cross-attention), so N <= FF_MAX_PASS - 1 -- checked
        here, before the first UNet call, instead of deep inside the first TCA layer (the reference loops over any N)."""
        from ._lib import FF_MAX_PASS
        n_src = len(ori_mask_lists)
        if n_src + 1 > FF_MAX_PASS:
            raise ValueError(f"cross-image composition with {n_src} source images needs {n_src + 1} attention passes per "
                             f"head; this build supports FF_MAX_PASS = {FF_MAX_PASS} (at most {FF_MAX_PASS - 1} sources)")
        init_code_orig = deepcopy(inverted_latents[-1])
        full_h, full_w = source_image.shape[:2]
        tgt_masks, ori_masks, local_pert, comp_cfg = self.prepare_composition_masks(
            ori_mask_lists, tgt_mask_lists, full_h, full_w, init_code_orig, dil_completion=dil_completion,
            dil_factor=dil_factor, draw_mask=draw_mask, appearance_transfer=appearance_transfer)
        c = self.controller
        c.src_masks, c.tgt_masks = ori_masks.to(self.device), tgt_masks.to(self.device)
        c.reset()
        image, intermediates = self.forward_sampling_compose(
            prompt=edit_prompt_list, refer_latents=inverted_latents[::-1], end_step=end_step,
            batch_size=init_code_orig.shape[0], latents=init_code_orig, guidance_scale=guidance_scale,
            num_inference_steps=num_steps, num_actual_inference_steps=num_steps - start_step, eta=eta,
            local_var_reg=local_pert, local_edit_text=local_text_edit, cfg_masks_tensor=comp_cfg, share_attn=share_attn,
            method_type=method_type, verbose=verbose, local_perturbation=local_perturbation,
            return_intermediates=return_intermediates, use_share_attention=use_share_attention, end_scale=end_scale)
        c.reset()
        return (image.permute(1, 2, 0).detach().cpu().numpy() * 255).astype(np.uint8), intermediates

    def FreeFine_cross_image_composition(self, img_lists, ori_mask_lists, tgt_mask_lists, coarse_input,
                                         guidance_text_list, guidance_scale, eta, end_step=10, num_step=50,
                                         start_step=25, share_attn=True, method_type='tca', local_text_edit=True,
                                         local_perturbation=True, verbose=True, seed=42, draw_mask=None,
                                         return_intermediates=False, use_auto_draw=False, end_scale=0.5,
                                         dil_completion=False, dil_factor=15, appearance_transfer=False):
        """reference model.py:1051-1086 (needs register_attention_control_compose).  As published the reference raises
        TypeError here -- it forwards `use_auto_draw` to a function without that parameter; this
        entry point accepts the argument and drops it, the inner functions are the parity targets."""
        methods = ['tca', 'ssa', 'sdsa', 'mmsa', 'mmsa_es']
        assert method_type in methods, f"check method type f{method_type}, which is not in {methods}"
        seed_everything(seed)
        ori_mask_lists = [self.mask_reduce_dim(m) for m in ori_mask_lists]
        tgt_mask_lists = [self.mask_reduce_dim(m) for m in tgt_mask_lists]
        inverted = self.DDIM_inversion_func_compose(img=coarse_input, compose_imgs=img_lists, prompt="", num_step=num_step,
                                                    start_step=start_step, verbose=verbose)
        image, intermediates = self.Details_Preserving_regeneration_compose(
            coarse_input, inverted, guidance_text_list, ori_mask_lists, tgt_mask_lists, draw_mask, num_steps=num_step,
            start_step=start_step, end_step=end_step, dil_factor=dil_factor, guidance_scale=guidance_scale, eta=eta,
            share_attn=share_attn, method_type=method_type, verbose=verbose, dil_completion=dil_completion,
            local_text_edit=local_text_edit, local_perturbation=local_perturbation,
            return_intermediates=return_intermediates, end_scale=end_scale, appearance_transfer=appearance_transfer)
        self.last_intermediates = intermediates
        return image

    # ------------------------------------------------------------------------------------------------------------
    # batched entry point (extension: E edits share one stream batch; the reference loops over edits one by one,
    # evaluation/FreeFine/freefine_batch_infer_2d.py:177-234)
    # ------------------------------------------------------------------------------------------------------------
    def prepare_various_mask_batch(self, shifted, ori, draw, cons, lat_hw, use_auto_draw, reduce_inp_artifacts):
        """prepare_various_mask (model.py:1432-1512) for E edits at once on device tensors [E,H,W] uint8 (any nonzero =
        set).  Same uint8 algebra (wrap-around included, quirk Q1) as the per-edit method."""
        if shifted.is_cuda:                     # one launch (ff_mask_prep); the algebra below is its CPU restatement
            need_cons = use_auto_draw or reduce_inp_artifacts
            return ops.mask_prep(shifted.contiguous(), ori.contiguous(), None if use_auto_draw else draw.contiguous(),
                                 cons.contiguous() if need_cons else None, lat_hw, use_auto_draw, reduce_inp_artifacts)
        b = lambda t: (t > 0).to(torch.uint8)

        def dil(t, k):
            a = k // 2
            x = F.pad(b(t)[:, None].float(), (a, k - 1 - a, a, k - 1 - a), value=0.0)
            return F.max_pool2d(x, kernel_size=k, stride=1)[:, 0].to(torch.uint8)

        sh, orim = b(shifted), b(ori)
        if not use_auto_draw:
            flex = b(draw) * (1 - sh)
            fg = flex + sh
            fg[fg > 0] = 1
            comp = flex
            if not reduce_inp_artifacts:
                lvar = flex
            else:
                lvar = (1 - b(cons)) * (1 - sh) * dil(ori, 30) + flex
                lvar[lvar > 0] = 1
        else:
            fg = sh
            c2 = b(cons) - orim
            if not reduce_inp_artifacts:
                comp = (1 - c2) * (1 - sh) * dil(shifted, 15)
            else:
                comp = dil(ori, 30) + dil(shifted, 15)
                comp[comp > 0] = 1
                comp = comp * ((1 - c2) * (1 - sh))
            lvar = comp
        ds = lambda t: F.interpolate(t[:, None], lat_hw, mode='nearest')[:, 0].contiguous()
        return fg, sh, orim, ds(comp), ds(lvar)

    @torch.no_grad()
    def FreeFine_generation_batch(self, ori_imgs, ori_masks, edit_params, guidance_texts, guidance_scale=7.5, eta=1.0,
                                  end_step=50, num_step=50, start_step=35, method_type='tca', seed=42, draw_masks=None,
                                  use_auto_draw=True, reduce_inp_artifacts=True, end_scale=0.0, inp_bgs=None,
                                  thetas=None, return_latents=False, coarse_inputs=None, target_masks=None):
        """E whole 2-D edits (coarse warp+blend -> DDIM inversion -> TCA sampling -> decode) in one stream batch.
        ori_imgs: uint8 [E,H,W,3] (numpy / pinned host tensor / CUDA tensor), ori_masks: uint8 [E,H,W] 0/1,
        edit_params: list of (dx,dy,rz,sx,sy) (or precomputed `thetas` f32 [E,2,3] when the masks live on the device).
        Defaults are the reference's GeoBench-2D settings (freefine_batch_infer_2d.py:212-230).
        coarse_inputs u8 [E,H,W,3] + target_masks u8 [E,H,W] (any non-zero = set): the coarse edit was made elsewhere
        (GeoBench-3D: the depth / novel-view pre-step, freefine_batch_infer_3d_depth.py:117-122) -- the warp is skipped,
        edit_params / thetas are ignored and cons_area = target mask as in that driver (:160).
        Returns uint8 images [E,H,W,3] on the side the inputs came from (host inputs -> host numpy)."""
        from . import coarse_edit
        assert method_type in ['tca', 'mmsa', 'mmsa_es'], method_type
        seed_everything(seed)
        host_in = not (torch.is_tensor(ori_imgs) and ori_imgs.is_cuda)
        to_dev = lambda a: (a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))).to(self.device, non_blocking=True)
        imgs_u8, masks = to_dev(ori_imgs), to_dev(ori_masks)
        E, H, W = masks.shape
        if (coarse_inputs is None) != (target_masks is None):
            raise ValueError("coarse_inputs and target_masks go together")
        if coarse_inputs is None and thetas is None:
            mh = ori_masks.cpu().numpy() if torch.is_tensor(ori_masks) else np.asarray(ori_masks)
            thetas = torch.tensor(np.stack([coarse_edit.cv2_theta(coarse_edit.edit_matrix(mh[e], edit_params[e]), W, H)
                                            for e in range(E)]), dtype=torch.float32)
        imgs = imgs_u8.permute(0, 3, 1, 2).float().contiguous()
        bgs = imgs if inp_bgs is None else to_dev(inp_bgs).permute(0, 3, 1, 2).float().contiguous()
        if coarse_inputs is not None:
            coarse = to_dev(coarse_inputs).permute(0, 3, 1, 2).float().contiguous()
            tgt = to_dev(target_masks).contiguous()
        else:
            with ops.nvtx_range("ff.coarse_edit (warp + blend)"):
                coarse, tgt = coarse_edit.re_edit_2d_device(imgs, masks.contiguous(), to_dev(thetas), bgs)
            coarse = coarse.round().clamp(0, 255)                               # the reference hands a uint8 image on
        draw = torch.zeros_like(masks) if draw_masks is None else to_dev(draw_masks)
        lat_hw = (H // 8, W // 8)
        fg, sh, orim, comp, lvar = self.prepare_various_mask_batch(tgt, masks, draw, tgt, lat_hw, use_auto_draw,
                                                                   reduce_inp_artifacts)
        src = torch.stack([coarse, imgs], dim=1).reshape(2 * E, 3, H, W) / 127.5 - 1
        _, inverted = self.invert(src, [""] * (2 * E), guidance_scale=1.0, num_inference_steps=num_step,
                                  num_actual_inference_steps=num_step - start_step, return_intermediates=True)
        c = self.controller
        c.reset()
        c.fg_retain_mask, c.fg_retain_mask_st2, c.fg_ref_mask, c.local_edit_region = fg, sh, orim, fg
        prompts = [p for t in guidance_texts for p in (t, "")]
        # the reference seeds every edit alike (seed_everything(seed), model.py:1018) and draws [2,C,h,w] per step: one
        # generator per edit reproduces that stream for each edit, independent of E, of its position in the batch and of
        # how a sweep is sharded over ranks
        gens = [torch.Generator(device=self.device).manual_seed(int(seed)) for _ in range(E)] if eta > 0 else None
        image, inter = self.forward_sampling(prompt=prompts, refer_latents=inverted[::-1], end_step=end_step, generator=gens,
                                             latents=inverted[-1].clone(), guidance_scale=guidance_scale,
                                             num_inference_steps=num_step, num_actual_inference_steps=num_step - start_step,
                                             eta=eta, completion_mask_cfg=comp, local_var_reg=lvar, method_type=method_type,
                                             end_scale=end_scale, return_intermediates=True, verbose=True)
        c.reset()
        out = (image.view(E, 2, 3, H, W)[:, 0].permute(0, 2, 3, 1) * 255).to(torch.uint8)
        if host_in:
            out = out.cpu().numpy()
        if return_latents:
            return out, inter[-1]
        return out


__all__ = ["FreeFinePipeline", "Attention_Modulator", "register_attention_control",
           "register_attention_control_4bggen", "register_attention_control_compose", "randn_tensor"]
