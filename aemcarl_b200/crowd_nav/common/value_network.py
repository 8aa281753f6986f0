"""ValueNetworkTransformerAndGRU mirror (crowd_nav/common/qnetwork_factory.py:586-687, components.py:8-137).

An nn.Module with the reference's parameter names and shapes (including the dead `encoder_layer.*` template), so
reference checkpoints load with `get_model().load_state_dict(torch.load(...))` and the trainer can optimise it.

forward(state[B, n, 13]) -> (value[B, 1], 0.5):
  * B == 1 without autograd (what MultiHumanRL.predict and Explorer.update_memory do)  -> CUDA kernel
    aem_value_forward (per-sample halting == the reference's batch-global halting at B = 1);
  * otherwise (training batches, trainer.py:52,75) -> a differentiable torch restatement that keeps the reference's
    BATCH-GLOBAL ACT halting (components.py:125-134).  This is the training path (autograd plumbing), not the
    measured hot path; the hot path is aem_lookahead, fed directly by ACTENVCARL.predict.
Modules are created in the reference's order, so torch.manual_seed(s) before construction yields the same
initial weights as the reference.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from aemcarl_b200 import weights as wts


def mlp(input_dim, mlp_dims, last_relu=False):
    layers = []
    dims = [input_dim] + list(mlp_dims)
    for i in range(len(dims) - 1):
        layers.append(nn.Linear(dims[i], dims[i + 1]))
        if i != len(dims) - 2 or last_relu:
            layers.append(nn.ReLU())
    return nn.Sequential(*layers)


class GRUEXND(nn.Module):
    """GRU cell whose update / reset gates are 2-layer MLPs ending in ReLU (components.py:38-53)."""

    def __init__(self, input_dim, hidden_dims, last_relu=True):
        super().__init__()
        self.convz = mlp(input_dim + hidden_dims[-1], hidden_dims, last_relu=last_relu)
        self.convr = mlp(input_dim + hidden_dims[-1], hidden_dims, last_relu=last_relu)
        self.convq = nn.Linear(hidden_dims[-1] + input_dim, hidden_dims[-1])

    def forward(self, x, h):
        hx = torch.cat([h, x], dim=1)
        z = torch.sigmoid(self.convz(hx))
        r = torch.sigmoid(self.convr(hx))
        q = torch.tanh(self.convq(torch.cat([r * h, x], dim=1)))
        return (1 - z) * h + z * q


class ATCBasic(nn.Module):
    """Adaptive computation time over the GRU cell (components.py:87-137)."""

    def __init__(self, input_dim, rnn_hidden_dims, max_ponder=3, epsilon=0.05, last_relu=True, act_steps=3,
                 act_fixed=False):
        super().__init__()
        self.rnn_hidden_dim = rnn_hidden_dims[-1]
        self.epsilon = epsilon
        self.rnn_cell = GRUEXND(input_dim, rnn_hidden_dims, last_relu)
        self.max_ponder = max_ponder
        self.ponder_linear = nn.Linear(rnn_hidden_dims[-1], 1)
        self.act_fixed = act_fixed
        self.act_steps = act_steps

    def forward(self, input, hx=None):
        x = input.reshape(-1, input.shape[2])
        if self.act_fixed:
            for _ in range(self.act_steps):
                hx = self.rnn_cell(x, hx)
            return hx, self.act_steps
        accum_p = 0
        accum_hx = torch.zeros([x.shape[0], self.rnn_hidden_dim], device=x.device, dtype=x.dtype)
        step_count = 0
        for _ in range(self.max_ponder):
            hx = self.rnn_cell(x, hx)
            halten_p = torch.sigmoid(self.ponder_linear(hx))
            accum_p = accum_p + halten_p
            accum_hx = accum_hx + halten_p * hx
            step_count += 1
            if not bool((accum_p < 1 - self.epsilon).any()):      # batch-global, as the reference
                break
        return accum_hx / step_count, step_count

    def forward_per_sample(self, input, hx=None):
        """The halting rule of batch-1 calls applied to every sample of a batch on its own: sample b goes on while any of
        its rows has accum_p < 1 - epsilon (what components.py:132-134 tests when the reference evaluates one state at a
        time, explorer.py:137 / multi_human_rl.py:351); the kernels' rule.  Returns (mean accumulated state, steps [B])."""
        B, T = input.shape[0], input.shape[1]
        x = input.reshape(-1, input.shape[2])
        if self.act_fixed:
            for _ in range(self.act_steps):
                hx = self.rnn_cell(x, hx)
            return hx, torch.full((B,), self.act_steps, device=x.device, dtype=torch.int64)
        accum_p = torch.zeros([x.shape[0], 1], device=x.device, dtype=x.dtype)
        accum_hx = torch.zeros([x.shape[0], self.rnn_hidden_dim], device=x.device, dtype=x.dtype)
        active = torch.ones((B,), device=x.device, dtype=torch.bool)
        count = torch.zeros((B,), device=x.device, dtype=torch.int64)
        for k in range(1, self.max_ponder + 1):
            hx = self.rnn_cell(x, hx)
            halten_p = torch.sigmoid(self.ponder_linear(hx))
            rows = active.repeat_interleave(T).unsqueeze(1).to(x.dtype)
            accum_p = accum_p + halten_p * rows
            accum_hx = accum_hx + halten_p * hx * rows
            count = torch.where(active, torch.full_like(count, k), count)
            active = active & (accum_p.view(B, T) < 1 - self.epsilon).any(dim=1)
            if not bool(active.any()):
                break
        return accum_hx / count.repeat_interleave(T).unsqueeze(1).to(x.dtype), count

    def forward_static(self, input, hx=None):
        """The same function without host-side control flow, so that a training step can be captured in a CUDA graph:
        all max_ponder steps are evaluated and the result of the step at which the reference's batch-global test
        (components.py:132-134) stops is selected with tensor masks.  step_count comes back as a 0-dim tensor."""
        x = input.reshape(-1, input.shape[2])
        if self.act_fixed:
            for _ in range(self.act_steps):
                hx = self.rnn_cell(x, hx)
            return hx, self.act_steps
        accum_p = torch.zeros([x.shape[0], 1], device=x.device, dtype=x.dtype)
        accum_hx = torch.zeros([x.shape[0], self.rnn_hidden_dim], device=x.device, dtype=x.dtype)
        out = torch.zeros_like(accum_hx)
        count = torch.zeros((), device=x.device, dtype=torch.int64)
        stopped = torch.zeros((), device=x.device, dtype=torch.bool)
        for k in range(1, self.max_ponder + 1):
            hx = self.rnn_cell(x, hx)
            halten_p = torch.sigmoid(self.ponder_linear(hx))
            accum_p = accum_p + halten_p
            accum_hx = accum_hx + halten_p * hx
            stop_now = ~((accum_p < 1 - self.epsilon).any()) if k < self.max_ponder else torch.ones_like(stopped)
            take = stop_now & ~stopped
            out = torch.where(take, accum_hx / k, out)
            count = torch.where(take, torch.full_like(count, k), count)
            stopped = stopped | stop_now
        return out, count


class ValueNetworkTransformerAndGRU(nn.Module):
    def __init__(self, input_dim=13, self_state_dim=6, joint_state_dim=13, in_mlp_dims=(100, 50),
                 sort_mlp_dims=(100, 50), sort_mlp_attention=(50, 50, 1), action_dims=(100, 50, 1),
                 with_dynamic_net=True, with_global_state=False, multi_process_type="self_attention", act_steps=1,
                 act_fixed=False):
        super().__init__()
        if multi_process_type != "self_attention" or not with_dynamic_net:
            raise NotImplementedError("only --test_policy_flag 5 --multi_process self_attention runs in the reference "
                                      "(SURVEY.md hazard 6)")
        if list(in_mlp_dims) != [100, 50] or list(action_dims) != [100, 50, 1] or input_dim != 13:
            raise NotImplementedError("the CUDA kernels are specialised for the [actenvcarl] dims of policy.config")
        self.self_state_dim = self_state_dim
        self.in_mlp_dims = list(in_mlp_dims)
        self.input_dim = input_dim
        self.act_fixed = bool(act_fixed)
        self.act_steps = int(act_steps)
        self.multi_process_type = multi_process_type
        self.in_mlp = ATCBasic(input_dim, list(in_mlp_dims), epsilon=0.05, last_relu=True, act_steps=self.act_steps,
                               act_fixed=self.act_fixed)
        self.encoder_layer = nn.TransformerEncoderLayer(d_model=50, nhead=2, dim_feedforward=150)
        self.transformer_encoder = nn.TransformerEncoder(self.encoder_layer, num_layers=3,
                                                         enable_nested_tensor=False)
        self.action_mlp = mlp(50 + self_state_dim, list(action_dims))
        self.attention_weights = None
        self.step_cnt = 0
        self._engine = None
        self._loaded_version = None

    def get_step_cnt(self):
        return self.step_cnt

    def __deepcopy__(self, memo):
        """Explorer.update_target_model deep-copies the network (explorer.py:18-19); the copy must not share (or try to
        pickle) the ctypes engine handle -- it uploads its own weights on first use."""
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k in ("_engine", "_loaded_version") else copy.deepcopy(v, memo)
        return new

    # ---- flat blob <-> parameters ---------------------------------------------------------------
    def flat_weights(self):
        return wts.flatten_state_dict(self.state_dict())

    def load_flat_weights(self, flat):
        sd = self.state_dict()
        for k, v in wts.unflatten(flat).items():
            sd[k].copy_(torch.from_numpy(v.copy()))

    def note_external_update(self):
        """the parameters were updated through raw pointers (aem_adam_step): torch's version counters did not move"""
        self._extern_version = getattr(self, "_extern_version", 0) + 1

    def _param_version(self):
        return (getattr(self, "_extern_version", 0),) + tuple(p._version for p in self.parameters()) \
            + tuple(p.data_ptr() for p in self.parameters())

    def sync_engine(self, engine):
        """(re)upload the weights to `engine` when any parameter changed since the last upload"""
        ver = (id(engine), self._param_version())
        if self._loaded_version != ver:
            engine.load_weights(self.flat_weights())
            engine.set_act_mode(self.act_fixed, self.act_steps)
            self._loaded_version = ver
        self._engine = engine
        return engine

    # ---- forward --------------------------------------------------------------------------------
    def _robot_token(self, state):
        tok = state[:, 0, :].clone().detach()
        tok[:, 6] = 0.0
        tok[:, 7] = 0.0
        tok[:, 8] = tok[:, 4]
        tok[:, 9] = tok[:, 5]
        tok[:, 10] = tok[:, 3]
        tok[:, 11] = 0.0
        tok[:, 12] = tok[:, 3]
        return tok.unsqueeze(1)

    def forward_torch(self, state, static=False, per_sample=False):
        """differentiable restatement of qnetwork_factory.py:641-687 (dropout inactive: the module is used in
        eval mode -- see hazard 1 in SURVEY.md; call .train() to get the reference's stochastic behaviour).
        static=True: no host-side control flow (CUDA-graph capturable), same result.  per_sample=True: every sample of
        the batch halts on its own, i.e. the result of evaluating the states one at a time like explorer.py:137."""
        size = state.shape
        self_state = state[:, 0, :self.self_state_dim].clone().detach()
        x = torch.cat([self._robot_token(state), state], dim=1)
        h0 = torch.zeros([size[0] * (size[1] + 1), self.in_mlp_dims[-1]], device=state.device, dtype=state.dtype)
        if per_sample:
            out, self.step_cnt = self.in_mlp.forward_per_sample(x, h0)
        else:
            out, self.step_cnt = self.in_mlp.forward_static(x, h0) if static else self.in_mlp(x, h0)
        enc_in = out.view(size[0], -1, 50).transpose(0, 1).contiguous()
        enc_out = self.transformer_encoder(enc_in).transpose(0, 1).contiguous()
        env_info = enc_out[:, 0, :]
        value = self.action_mlp(torch.cat([self_state, env_info], dim=1))
        return value.view(size[0], -1), 0.5

    def forward(self, state):
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if state.is_cuda and state.shape[0] == 1 and not needs_grad and not self.training:
            from aemcarl_b200.engine import get_engine
            eng = self.sync_engine(self._engine or get_engine(state.device.index or 0, role="policy"))
            v, steps = eng.value_forward(state.detach().to(torch.float32).contiguous())
            self.step_cnt = int(steps.max().item())
            return v.view(1, 1), 0.5
        return self.forward_torch(state)
