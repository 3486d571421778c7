"""Shared machinery of the two reference models that train on (pos-pair, neg-pair) batches:
``CFL`` (dist half, cfl/models/cfl.py:683-728, 868-949, 1065-1085) and ``Dist``
(cfl/models/dist.py:214-293).

The TF graph + ``sess.run`` of the reference becomes an eager object: ``train_step`` does what
``sess.run([s_optim, update_stats, ...])`` does (forward of the pos and neg pair batches, loss,
backward, TF-style Adam) with the fused sm_100a kernels -- projection, paired distance + loss
forward, paired backward, projection backward, Adam -- and no autograd; ``predict`` is
``sess.run(val_s_pos_predicts.outputs)``.  Graph-node attribute names (``s_pos_dists``,
``s_accuracy`` ...) hold the tensors / scalars of the last executed step.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from .. import _native as nat
from .. import variables as vs
from .base import ModelBase

_ACT = {None: None, "linear": None, "tanh": "tanh", "sigmoid": "sigmoid", "relu": "relu"}


# tf.maximum(raw_threshold, 1e-6) compares in float32: theta is INITIALISED at float32(1e-6), which is
# below the double 1e-6, so the floor must be the float32 constant for the tie to pass the gradient
_THETA_FLOOR32 = float(torch.tensor(1e-6, dtype=torch.float32))


class Head:
    """One FC head: weight-normalised (g, V, biases) or plain (weights, biases)."""

    def __init__(self, fan_in, fan_out, weight_norm, bias, initializer=None):
        init = initializer or vs.xavier_initializer()
        with vs.variable_scope("fully_connected"):
            if weight_norm:
                self.g = vs.get_variable("g", [fan_out], vs.ones_initializer())
                self.V = vs.get_variable("V", [fan_in, fan_out], init)
            else:
                self.g = None
                self.V = vs.get_variable("weights", [fan_in, fan_out], init)
            self.b = vs.get_variable("biases", [fan_out], vs.zeros_initializer()) if bias else None
        self.weight_norm = weight_norm

    def params(self):
        return [p for p in (self.V, self.g, self.b) if p is not None]


class Encoder:
    """The variables of one encoder scope (``DistEncoder`` / ``DistEncoderSrc`` / ``Encoder``)."""

    def __init__(self, scope_name, F, d, K, dist_type, weight_norm, head_names, initializer=None):
        has_bias = dist_type.startswith("pcd") or not weight_norm       # base.py:45-46; dist.py:52,65
        with vs.variable_scope(scope_name) as sc:
            self.name = sc.name
            with vs.variable_scope(head_names[0]):
                self.e0 = Head(F, d, weight_norm, has_bias, initializer)
            self.proto = None
            if dist_type in ("pcd", "monomer"):
                with vs.variable_scope(head_names[1]):
                    self.proto = Head(F, K * d, weight_norm, has_bias, initializer)
            self.gate = None
            if dist_type == "monomer":
                with vs.variable_scope("monomer_outputs"):
                    self.gate = Head(d, K, True, False, initializer)

    def heads(self):
        return [h for h in (self.e0, self.proto, self.gate) if h is not None]


class PairModel(ModelBase):
    def _init_pair_model(self, *, input_size, latent_size, num_components, dist_type, act_type,
                         weight_norm, pos_weight, use_threshold, caffe_margin, lambda_m, reg_const,
                         directed, lr, beta1, beta2, in_scale, head_names, encoder_names, initializer=None):
        self.input_size, self.latent_size, self.num_components = input_size, latent_size, num_components
        self.dist_type, self.act = dist_type, _ACT[act_type]
        self.weight_norm = weight_norm
        self.pos_weight, self.use_threshold = pos_weight, use_threshold
        self.caffe_margin, self.lambda_m, self.reg_const = caffe_margin, lambda_m, reg_const
        self.directed = directed
        self.lr, self.beta1, self.beta2 = lr, beta1, beta2
        self.in_scale = in_scale
        K = num_components if dist_type in ("pcd", "monomer") else 1
        self.K = K
        self.enc_src = Encoder(encoder_names[0], input_size, latent_size, K, dist_type, weight_norm,
                               head_names, initializer)
        self.enc_dst = self.enc_src if not directed else Encoder(
            encoder_names[1], input_size, latent_size, K, dist_type, weight_norm, head_names, initializer)
        with vs.variable_scope("Thresholder"):
            with vs.variable_scope("threshold"):
                self.raw_threshold = vs.get_variable("threshold", [], vs.constant_initializer(1e-6))
        self._params = []
        for enc in ([self.enc_src] if not directed else [self.enc_src, self.enc_dst]):
            for h in enc.heads():
                self._params.extend(h.params())
        self.s_vars = list(self._params)
        self.th_vars = [self.raw_threshold]
        self._adam = {id(p): (torch.zeros_like(p), torch.zeros_like(p)) for p in self._params + self.th_vars}
        self._grads = {id(p): torch.zeros_like(p) for p in self._params + self.th_vars}
        self._step = 0
        self._th_step = 0
        self._ema = {}
        self.ema_decay = 0.99

    # ------------------------------------------------------------------------------------------
    @property
    def threshold(self):
        return torch.clamp(self.raw_threshold.detach(), min=1e-6)

    def _encode(self, xs, xt, want_bwd):
        """Projects one (source, target) pair batch; returns the pair-kernel operands + saved state."""
        S, D = self.enc_src, self.enc_dst
        sc, act, dt = self.in_scale, self.act, self.dist_type
        st = SimpleNamespace(xs=xs, xt=xt)
        if dt.startswith("pcd"):
            h = S.proto
            st.P, _, st.zP = nat.project_fwd(xs, h.V, h.g, h.b, h.weight_norm, sc, act, want_z=want_bwd and h.weight_norm)
            h = D.e0
            st.a, _, st.za = nat.project_fwd(xt, h.V, h.g, h.b, h.weight_norm, sc, act, want_z=want_bwd and h.weight_norm)
            st.w = None
            st.mode = "pcd"
        elif dt == "siamese":
            h = S.e0
            st.a, _, st.za = nat.project_fwd(xs, h.V, h.g, h.b, h.weight_norm, sc, act, want_z=want_bwd and h.weight_norm)
            h = D.e0
            st.P, _, st.zP = nat.project_fwd(xt, h.V, h.g, h.b, h.weight_norm, sc, act, want_z=want_bwd and h.weight_norm)
            st.w = None
            st.mode = "siamese"
        else:                                                   # monomer (base.py:94-117)
            h = S.e0
            st.a, st.pre, st.za = nat.project_fwd(xs, h.V, h.g, h.b, h.weight_norm, sc, act, want_pre=True, want_z=want_bwd)
            g = S.gate
            st.logits, _, st.zg = nat.project_fwd(st.pre, g.V, g.g, None, True, 1.0, None, want_z=want_bwd)
            st.w = torch.softmax(st.logits, dim=-1)
            h = D.proto
            st.P, _, st.zP = nat.project_fwd(xt, h.V, h.g, h.b, h.weight_norm, sc, act, want_z=want_bwd and h.weight_norm)
            st.mode = "monomer"
        st.P3 = st.P.view(-1, self.K, self.latent_size)
        return st

    def _pair_fwd(self, st, label):
        margin = self.caffe_margin or 0.0
        st.dist, st.score, _, st.stats = nat.pair_loss_fwd(
            st.mode, st.a, st.P3, w=st.w, theta=self.raw_threshold, label=label, margin=margin,
            want_score=True, want_stats=True)
        return st

    def predict(self, source, target):
        """``sess.run(model.val_s_pos_predicts.outputs, feed)``: scores [B,1] = theta+ - dist."""
        with torch.no_grad():
            st = self._encode(self._prep(source), self._prep(target), want_bwd=False)
            _, score, _, _ = nat.pair_loss_fwd(st.mode, st.a, st.P3, w=st.w, theta=self.raw_threshold,
                                               want_dist=False, want_score=True)
        return score.view(-1, 1)

    def dists(self, source, target):
        with torch.no_grad():
            st = self._encode(self._prep(source), self._prep(target), want_bwd=False)
            dist, _, _, _ = nat.pair_loss_fwd(st.mode, st.a, st.P3, w=st.w)
        return dist.view(-1, 1)

    def _prep(self, x):
        x = torch.as_tensor(x)
        if not x.is_cuda:
            x = x.to(vs.default_device(), non_blocking=True)
        return x.reshape(x.shape[0], -1).float().contiguous()

    # ------------------------------------------------------------------------------------------
    def _head_bwd(self, head, x, y, z, dy, act, first, in_scale):
        gV, gg, gb = self._grads[id(head.V)], (self._grads[id(head.g)] if head.g is not None else None), \
            (self._grads[id(head.b)] if head.b is not None else None)
        acc = not first[id(head.V)]
        nat.project_bwd(x, head.V, head.g, head.b, head.weight_norm, in_scale, act, y, z, dy,
                        dV=gV, dg=gg, dbias=gb, accumulate=acc,
                        reg_c=(self.reg_const if not acc else 0.0), want_dg=gg is not None, want_dbias=gb is not None)
        first[id(head.V)] = False

    def _pair_bwd(self, st, label, B, first, want_ce):
        pw = self.pos_weight if self.pos_weight else 1.0
        use_ce = self.use_threshold and want_ce
        if label == 1:
            c_ce = (pw / B) if use_ce else 0.0
            c_lin = (0.5 * pw / B) if self.caffe_margin else ((self.lambda_m * pw / B) if self.lambda_m else 0.0)
            c_mar = 0.0
        else:
            c_ce = (1.0 / B) if use_ce else 0.0
            c_lin = 0.0
            c_mar = (0.5 / B) if self.caffe_margin else 0.0
        da, dP, dw, dth = nat.pair_loss_bwd(st.mode, st.a, st.P3, w=st.w, theta=self.raw_threshold, label=label,
                                            margin=self.caffe_margin or 0.0, c_ce=c_ce, c_lin=c_lin,
                                            c_margin=c_mar, want_dtheta=True)
        S, D, sc, act = self.enc_src, self.enc_dst, self.in_scale, self.act
        dP2 = dP.reshape(dP.shape[0], -1)
        if st.mode == "pcd":
            self._head_bwd(S.proto, st.xs, st.P, st.zP, dP2, act, first, sc)
            self._head_bwd(D.e0, st.xt, st.a, st.za, da, act, first, sc)
        elif st.mode == "siamese":
            self._head_bwd(S.e0, st.xs, st.a, st.za, da, act, first, sc)
            self._head_bwd(D.e0, st.xt, st.P, st.zP, dP2, act, first, sc)
        else:
            # gate: dw -> softmax -> weight-norm FC on the pre-activation e0 (tiny [B,K] plumbing)
            dlog = st.w * (dw - (dw * st.w).sum(-1, keepdim=True))
            g = S.gate
            self._head_bwd(g, st.pre, None, st.zg, dlog, None, first, 1.0)
            sg = (g.g / g.V.detach().pow(2).sum(0).sqrt())
            dpre = dlog @ (g.V.detach() * sg).t()
            if act is not None:
                from ..functional import _act_grad
                dpre = dpre + da * _act_grad(st.a, act)
            else:
                dpre = dpre + da
            self._head_bwd(S.e0, st.xs, None, st.za, dpre.contiguous(), None, first, sc)
            self._head_bwd(D.proto, st.xt, st.P, st.zP, dP2, act, first, sc)
        return dth

    def train_step(self, src_pos, dst_pos, src_neg, dst_neg, val_batches=None):
        """One optimiser step on a (pos pairs, neg pairs) batch -- the reference's hot loop
        ``sess.run([summary, [s_optim, update_stats], ...])`` (cfl/models/cfl.py:1405-1414,
        cfl/bin/train_dist.py:81-82).  Returns a dict of the fetched scalars."""
        xsp, xtp, xsn, xtn = (self._prep(t) for t in (src_pos, dst_pos, src_neg, dst_neg))
        Bp, Bn = xsp.shape[0], xsn.shape[0]
        sp = self._pair_fwd(self._encode(xsp, xtp, True), 1)
        sn = self._pair_fwd(self._encode(xsn, xtn, True), 0)
        first = {id(p): True for p in self._params}
        dth_p = self._pair_bwd(sp, 1, Bp, first, want_ce=True)
        dth_n = self._pair_bwd(sn, 0, Bn, first, want_ce=True)
        world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            world = torch.distributed.get_world_size()
        # one device->host read per step for the scalars the reference fetches every step
        host = torch.cat([sp.stats, sn.stats, dth_p, dth_n]).cpu().tolist()
        stp, stn, dthp, dthn = host[0:8], host[8:16], host[16], host[17]
        out = self._losses_from_stats(stp, stn, Bp, Bn)
        # theta gradient: -(sum CE part) with the tf.maximum tie rule (blocks.py:20-21)
        th_live = float(self.raw_threshold.detach()) >= _THETA_FLOOR32
        th_grad = (dthp + dthn) if th_live else 0.0
        if not self.use_threshold:
            # theta is trained by its own Adam on s_thres_loss (cfl.py:1076-1079): recompute the CE sums
            pw = self.pos_weight if self.pos_weight else 1.0
            _, _, _, a_ = nat.pair_loss_bwd(sp.mode, sp.a, sp.P3, w=sp.w, theta=self.raw_threshold, label=1,
                                            c_ce=pw / Bp, want_dtheta=True)
            _, _, _, b_ = nat.pair_loss_bwd(sn.mode, sn.a, sn.P3, w=sn.w, theta=self.raw_threshold, label=0,
                                            c_ce=1.0 / Bn, want_dtheta=True)
            th_grad = float(a_ + b_) if th_live else 0.0
        self._grads[id(self.raw_threshold)].fill_(th_grad)
        # heads the loss never reads (directed models: e.g. DistEncoderSrc/outputs under pcd): TF gives
        # their g no gradient (Adam skips it) and their V / biases only the regulariser's c*w
        idle = set()
        for enc in ([self.enc_src] if not self.directed else [self.enc_src, self.enc_dst]):
            for h in enc.heads():
                if not first[id(h.V)]:
                    continue
                for p_ in h.params():
                    if self.reg_const and p_ is not h.g:
                        self._grads[id(p_)].copy_(p_.detach()).mul_(self.reg_const)
                    else:
                        self._grads[id(p_)].zero_()
                        idle.add(id(p_))
        if world > 1:
            flat = torch.cat([self._grads[id(p)].reshape(-1) for p in self._params + self.th_vars])
            torch.distributed.all_reduce(flat)
            o = 0
            for p in self._params + self.th_vars:
                n = p.numel()
                self._grads[id(p)].copy_(flat[o:o + n].view_as(p))
                o += n
        self._step += 1
        for p in self._params + self.th_vars:
            if id(p) in idle:
                continue
            m, v = self._adam[id(p)]
            nat.adam_step(p.data.view(-1), self._grads[id(p)].view(-1), m.view(-1), v.view(-1), self._step,
                          self.lr, self.beta1, self.beta2, 1e-8, 1.0 / world)
        # graph-node attributes of the reference
        self.s_pos_dists, self.s_neg_dists = sp.dist.view(-1, 1), sn.dist.view(-1, 1)
        self.s_pos_predicts = SimpleNamespace(outputs=sp.score.view(-1, 1), threshold=self.threshold)
        self.s_neg_predicts = SimpleNamespace(outputs=sn.score.view(-1, 1), threshold=self.threshold)
        for k, v_ in out.items():
            setattr(self, k, v_)
        self._update_ema(out)
        if val_batches is not None:
            out["val_s_accuracy"] = self.val_accuracy(*val_batches)
            self.val_s_accuracy = out["val_s_accuracy"]
            self._ema_one("val_s_accuracy", out["val_s_accuracy"])
        return out

    # ---- the same step with no host round trip, replayed as ONE CUDA graph --------------------------
    _SCALARS = ("s_p_loss_pos", "s_p_loss_neg", "s_thres_loss", "s_cd_loss", "s_loss_reg", "s_total_loss",
                "s_accuracy", "s_margins", "s_pos_dists_adapt", "s_neg_dists_adapt", "s_margin_adapt",
                "val_s_accuracy")
    _EMA_KEYS = ("s_accuracy", "s_margin_adapt", "s_pos_dists_adapt", "s_neg_dists_adapt", "val_s_accuracy")

    def _device_step(self, xsp, xtp, xsn, xtn, val=None):
        """train_step with every decision taken on the device: theta's tf.maximum gate, Adam's step count
        (cfl_adam_step_dev), the fetched scalars and their moving averages stay in device tensors
        (``fetch_scalars`` reads them).  ~60 small launches: captured once, replayed per batch."""
        Bp, Bn = xsp.shape[0], xsn.shape[0]
        sp = self._pair_fwd(self._encode(xsp, xtp, True), 1)
        sn = self._pair_fwd(self._encode(xsn, xtn, True), 0)
        first = {id(p): True for p in self._params}
        dth_p = self._pair_bwd(sp, 1, Bp, first, want_ce=True)
        dth_n = self._pair_bwd(sn, 0, Bn, first, want_ce=True)
        th_sum = dth_p + dth_n
        if not self.use_threshold:
            pw = self.pos_weight if self.pos_weight else 1.0
            _, _, _, a_ = nat.pair_loss_bwd(sp.mode, sp.a, sp.P3, w=sp.w, theta=self.raw_threshold, label=1,
                                            c_ce=pw / Bp, want_dtheta=True)
            _, _, _, b_ = nat.pair_loss_bwd(sn.mode, sn.a, sn.P3, w=sn.w, theta=self.raw_threshold, label=0,
                                            c_ce=1.0 / Bn, want_dtheta=True)
            th_sum = a_ + b_
        live = self.raw_threshold.detach() >= _THETA_FLOOR32
        self._grads[id(self.raw_threshold)].copy_(
            torch.where(live, th_sum.reshape(()).to(torch.float32), torch.zeros((), device=th_sum.device)))
        idle = set()
        for enc in ([self.enc_src] if not self.directed else [self.enc_src, self.enc_dst]):
            for h in enc.heads():
                if not first[id(h.V)]:
                    continue
                for p_ in h.params():
                    if self.reg_const and p_ is not h.g:
                        self._grads[id(p_)].copy_(p_.detach()).mul_(self.reg_const)
                    else:
                        self._grads[id(p_)].zero_()
                        idle.add(id(p_))
        reg = self.reg_loss_value()                          # of the weights the losses were computed with
        self._step_dev.add_(1)
        for p in self._params + self.th_vars:
            if id(p) in idle:
                continue
            m, v = self._adam[id(p)]
            nat.adam_step_dev(p.data.view(-1), self._grads[id(p)].view(-1), m.view(-1), v.view(-1), self._step_dev,
                              self.lr, self.beta1, self.beta2, 1e-8, 1.0)
        # the scalars of _losses_from_stats, on the device (float64)
        stp, stn = sp.stats, sn.stats
        pw = self.pos_weight
        lp, ln = stp[0] / Bp, stn[0] / Bn
        thres = lp * pw + ln if pw else lp + ln
        reg = reg.double() if torch.is_tensor(reg) else torch.zeros((), dtype=torch.float64, device=stp.device)
        total = reg + (thres if self.use_threshold else 0.0)
        mean_dp, mean_dn = stp[2] / Bp, stn[2] / Bn
        cd = torch.zeros((), dtype=torch.float64, device=stp.device)
        if self.caffe_margin:
            cd = 0.5 * ((mean_dp * pw if pw else mean_dp) + stn[4] / Bn)
            total = total + cd
        elif self.lambda_m:
            cd = mean_dp * self.lambda_m * (pw if pw else 1.0)
            total = total + cd
        acc = 0.5 * (stp[1] / Bp + stn[1] / Bn)
        pda, nda = stp[3] / Bp, stn[3] / Bn
        if val is not None:
            vp, vn = self.predict(val[0], val[1]), self.predict(val[2], val[3])
            vacc = 0.5 * ((vp > 0).double().mean() + (vn <= 0).double().mean())
        else:
            vacc = torch.zeros((), dtype=torch.float64, device=stp.device)
        vals = torch.stack([lp, ln, thres, cd, reg, total, acc, mean_dp - mean_dn, pda, nda, 0.5 * (pda + nda), vacc])
        self._dev_scalars.copy_(vals)
        sel = torch.stack([acc, 0.5 * (pda + nda), pda, nda, vacc])
        self._dev_ema.mul_(self.ema_decay).add_(sel, alpha=1.0 - self.ema_decay)
        self.s_pos_dists, self.s_neg_dists = sp.dist.view(-1, 1), sn.dist.view(-1, 1)
        self.s_pos_predicts = SimpleNamespace(outputs=sp.score.view(-1, 1), threshold=self.threshold)
        self.s_neg_predicts = SimpleNamespace(outputs=sn.score.view(-1, 1), threshold=self.threshold)

    def train_step_graph(self, src_pos, dst_pos, src_neg, dst_neg, val_batches=None):
        """One optimiser step replayed from a CUDA graph (single process only).  The first two calls of a
        given batch shape run the same device-only step eagerly (they warm every workspace), the third
        captures, later ones copy the batch into the graph's static inputs and replay.  Returns nothing:
        ``fetch_scalars()`` reads the last step's scalars and moving averages when they are wanted."""
        if torch.distributed.is_available() and torch.distributed.is_initialized() and \
                torch.distributed.get_world_size() > 1:
            raise RuntimeError("train_step_graph is single-process; use train_step for data-parallel training")
        xs = [self._prep(t) for t in (src_pos, dst_pos, src_neg, dst_neg)]
        vx = [self._prep(t) for t in val_batches] if val_batches is not None else None
        dev = xs[0].device
        if not hasattr(self, "_step_dev"):
            self._step_dev = torch.tensor(self._step, dtype=torch.int32, device=dev)
            self._dev_scalars = torch.zeros(len(self._SCALARS), dtype=torch.float64, device=dev)
            self._dev_ema = torch.zeros(len(self._EMA_KEYS), dtype=torch.float64, device=dev)
            self._graph, self._graph_key, self._graph_warm = None, None, 0
        key = (tuple(tuple(x.shape) for x in xs), None if vx is None else tuple(tuple(x.shape) for x in vx))
        if key != self._graph_key:
            self._graph, self._graph_key, self._graph_warm = None, key, 0
        if self._graph is None and self._graph_warm < 2:
            self._step_dev.fill_(self._step)
            self._device_step(*xs, val=vx)
            self._graph_warm += 1
        else:
            if self._graph is None:
                self._static_in = [torch.empty_like(x) for x in xs]
                self._static_val = [torch.empty_like(x) for x in vx] if vx is not None else None
                torch.cuda.synchronize()
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):
                    self._device_step(*self._static_in, val=self._static_val)
            for s_, x in zip(self._static_in, xs):
                s_.copy_(x)
            if vx is not None:
                for s_, x in zip(self._static_val, vx):
                    s_.copy_(x)
            self._graph.replay()
        self._step += 1

    def fetch_scalars(self):
        """Scalars of the last graph step (one device->host read), attributes and averages as train_step sets."""
        vals = self._dev_scalars.cpu().tolist()
        ema = self._dev_ema.cpu().tolist()
        out = dict(zip(self._SCALARS, vals))
        for k, v_ in out.items():
            setattr(self, k, v_)
        for k, b in zip(self._EMA_KEYS, ema):
            setattr(self, k + "_avg", b / (1 - self.ema_decay ** max(self._step, 1)))
        return out

    def val_accuracy(self, src_pos, dst_pos, src_neg, dst_neg):
        """cfl.py:941-949: accuracy of the val batch evaluated alongside each train step."""
        sp, sn = self.predict(src_pos, dst_pos), self.predict(src_neg, dst_neg)
        return 0.5 * (float((sp > 0).float().mean()) + float((sn <= 0).float().mean()))

    def _losses_from_stats(self, stp, stn, Bp, Bn):
        """_build_dist_losses (cfl.py:868-937) from the kernels' batch sums."""
        pw = self.pos_weight
        lp, ln = stp[0] / Bp, stn[0] / Bn
        thres = lp * pw + ln if pw else lp + ln
        reg = float(self.reg_loss_value())
        total = reg + (thres if self.use_threshold else 0.0)
        cd = 0.0
        mean_dp, mean_dn = stp[2] / Bp, stn[2] / Bn
        if self.caffe_margin:
            cd = 0.5 * ((mean_dp * pw if pw else mean_dp) + stn[4] / Bn)
            total += cd
        elif self.lambda_m:
            cd = mean_dp * self.lambda_m * (pw if pw else 1.0)
            total += cd
        return dict(s_p_loss_pos=lp, s_p_loss_neg=ln, s_thres_loss=thres, s_cd_loss=cd, s_loss_reg=reg,
                    s_total_loss=total, s_accuracy=0.5 * (stp[1] / Bp + stn[1] / Bn),
                    s_margins=mean_dp - mean_dn, s_pos_dists_adapt=stp[3] / Bp, s_neg_dists_adapt=stn[3] / Bn,
                    s_margin_adapt=0.5 * (stp[3] / Bp + stn[3] / Bn))

    def reg_loss_value(self):
        if not self.reg_const:
            return 0.0
        tot = 0.0
        encs = [self.enc_src] if not self.directed else [self.enc_src, self.enc_dst]
        for enc in encs:
            for h in enc.heads():
                tot = tot + 0.5 * self.reg_const * (h.V.detach() ** 2).sum()
                if h.b is not None:
                    tot = tot + 0.5 * self.reg_const * (h.b.detach() ** 2).sum()
        return tot

    def _ema_one(self, k, value):
        b = self._ema.get(k, 0.0) * self.ema_decay + (1 - self.ema_decay) * value
        self._ema[k] = b
        setattr(self, k + "_avg", b / (1 - self.ema_decay ** self._step))

    def _update_ema(self, out):
        """tf.train.ExponentialMovingAverage(0.99) over tensors (zero-debiased), cfl.py:528,903-949."""
        for k in ("s_accuracy", "s_margin_adapt", "s_pos_dists_adapt", "s_neg_dists_adapt"):
            self._ema_one(k, out[k])

    # ------------------------------------------------------------------------------------------
    def train(self, sess, data, start_iter, epochs, post_epochs, best_dir, best_acc_dir, checkpoint_dir,
              epoch_callback=None, post_epoch_callback=None, save_epochs=1, eval_epochs=1, save_iters=None,
              disable_eval=False, saver=None, best_saver=None, best_acc_saver=None, writer=None, check=None,
              cuda_graph=False):
        """The epoch loop of cfl/models/cfl.py:1349-1511 for the distance model (post epochs belong to
        the GAN half and are always 0 here, as in the reference when ``gan`` is off): per step one
        labelled train batch + one val batch; per ``eval_epochs`` a full ``dist_eval`` on val, and on a
        new best val AUC / best val accuracy a ``best_model`` / ``best_acc_model`` checkpoint plus the
        ``best_accuracy`` / ``best_accuracy_by_th`` stats files (written exactly as the reference
        writes them, including that both files carry the best-AUC triple)."""
        import logging
        import os
        from ..utils import dist_eval, load_best_stats
        log = logging.getLogger(__name__)
        nb_train = max(data.train.num_examples_labeled_pos, data.train.num_examples_labeled_neg)
        log.warning("%d pairs / %d images", nb_train, data.train.num_examples)
        nb_batch = nb_train // self.batch_size
        log.warning("%d batches per epoch", nb_batch)
        if nb_batch == 0:
            raise ValueError("batch_size %d exceeds the %d labelled training pairs" % (self.batch_size, nb_train))
        best_auc_path = os.path.join(best_dir, "best_accuracy")
        best_acc_path = os.path.join(best_acc_dir, "best_accuracy_by_th")
        stats, stats_acc = load_best_stats(best_auc_path), load_best_stats(best_acc_path)
        if sess is not None and getattr(sess, "model", None) is None:
            sess.model = self
        start_epoch = start_iter // nb_batch
        log.warning("start epoch %d of %d", start_epoch, epochs)
        train_avg = val_avg = 0.0
        for e in range(start_epoch, epochs):
            first = start_iter % nb_batch if e == start_epoch else 0
            for i in range(first, nb_batch):
                if cuda_graph:                               # no host round trip per step; scalars read per epoch
                    self.train_step_graph(*data.train.next_batch(self.batch_size),
                                          val_batches=data.val.next_batch(self.batch_size))
                    if i + 1 < nb_batch and not (save_iters and i > 0 and i % save_iters == 0):
                        continue
                    out = self.fetch_scalars()
                else:
                    out = self.train_step(*data.train.next_batch(self.batch_size),
                                          val_batches=data.val.next_batch(self.batch_size))
                train_avg, val_avg = self.s_accuracy_avg, self.val_s_accuracy_avg
                if writer is not None:
                    writer.add_summary(out, nb_batch * e + i)
                if save_iters and i > 0 and i % save_iters == 0 and saver is not None:
                    saver.save(sess, os.path.join(checkpoint_dir, "model"), global_step=nb_batch * e + i)
            if e % eval_epochs == 0 and not disable_eval:
                val_stats = dist_eval(sess, self, self.batch_size, data.val)
                if val_stats.auc > stats.best_auc or val_stats.accuracy > stats_acc.best_accuracy:
                    test_stats = dist_eval(sess, self, self.batch_size, data.test)
                    log.warning("epoch %d: current error = train: %f val: %f test: %f / auc = val: %f test: %f", e,
                                1. - train_avg, 1. - val_stats.accuracy, 1. - test_stats.accuracy, val_stats.auc,
                                test_stats.auc)
                    if val_stats.auc > stats.best_auc:
                        stats.best_accuracy, stats.best_auc, stats.best_epoch = val_stats.accuracy, val_stats.auc, e
                        if best_saver is not None:
                            best_saver.save(sess, os.path.join(best_dir, "model"), global_step=stats.best_epoch)
                        with open(best_auc_path, "w") as outfile:
                            outfile.write("{}\t{}\t{}".format(stats.best_epoch, stats.best_accuracy, stats.best_auc))
                    if val_stats.accuracy > stats_acc.best_accuracy:
                        stats_acc.best_accuracy, stats_acc.best_auc, stats_acc.best_epoch = \
                            val_stats.accuracy, val_stats.auc, e
                        if best_acc_saver is not None:
                            best_acc_saver.save(sess, os.path.join(best_acc_dir, "model"),
                                                global_step=stats.best_epoch if stats.best_epoch is not None else e)
                        with open(best_acc_path, "w") as outfile:
                            outfile.write("{}\t{}\t{}".format(stats.best_epoch, stats.best_accuracy, stats.best_auc))
                else:
                    log.warning("epoch %d: current error = train: %f val: %f / auc = val: %f", e, 1. - train_avg,
                                1. - val_stats.accuracy, val_stats.auc)
            else:
                log.warning("epoch %d: avg error = train: %f val: %f", e, 1. - train_avg, 1. - val_avg)
            if epoch_callback is not None:
                epoch_callback(e)
            if e % save_epochs == 0 and saver is not None:
                saver.save(sess, os.path.join(checkpoint_dir, "model"), global_step=(e + 1) * nb_batch)

    # ------------------------------------------------------------------------------------------
    def state_dict(self):
        """Variables by their reference names + optimiser slots."""
        named = vs.get_collection(self.name)
        sd = {k: v.detach().clone() for k, v in named.items()}
        sd["__step__"] = torch.tensor(self._step)
        for k, v in named.items():
            if id(v) in self._adam:
                sd[k + "/Adam"], sd[k + "/Adam_1"] = (t.clone() for t in self._adam[id(v)])
        return sd

    def load_state_dict(self, sd):
        named = vs.get_collection(self.name)
        for k, v in named.items():
            if k in sd:
                v.data.copy_(sd[k].to(v.device))
            if k + "/Adam" in sd and id(v) in self._adam:
                self._adam[id(v)][0].copy_(sd[k + "/Adam"].to(v.device))
                self._adam[id(v)][1].copy_(sd[k + "/Adam_1"].to(v.device))
        self._step = int(sd.get("__step__", 0))
