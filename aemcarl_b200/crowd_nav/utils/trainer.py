"""Value-network regression (crowd_nav/utils/trainer.py:22-86): MSE(model(inputs)[0], values) with SGD / Adam /
RMSprop exactly as configured in the reference.  The one addition is the data-parallel hook: when
torch.distributed is initialised, the flat gradient (113,752 live floats, one bucket) is all-reduced (NCCL over
NVLink on GPUs, gloo in the CPU tests) and averaged between loss.backward() and optimizer.step() -- the only
collective of the whole system.  Dropout: the reference never calls eval(), so it optimises with the four dropout
sites of every encoder layer live (p = 0.1); here the module's mode decides -- model.train() gives that regime in all three
implementations (torch's masks in the eager step, counter-based masks in the fused kernels), model.eval() (what
ACTENVCARL.configure leaves, because the lookahead kernels have no dropout) trains without it; train.py --dropout selects
the reference's regime.  Three implementations of one optimisation step, same arithmetic contract:
eager torch (the reference's code path), GraphedStep (the torch step replayed from CUDA graphs) and FusedStep (the hand-written
kernels of csrc/train.cu on one flat parameter vector: aem_train_grad + all-reduce + aem_adam_step / aem_sgd_step /
aem_rmsprop_step), selected by Trainer(cuda_graph=..., fused=...).  With a DeviceReplayMemory the mini-batches come from
aem_sample_batch."""
import logging

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.optim as optim
from torch.utils.data import DataLoader


def allreduce_gradients(model):
    """average gradients over ranks with ONE flat all-reduce; no-op without an initialised process group"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    grads = [p.grad for p in model.parameters() if p.grad is not None]
    if not grads:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= dist.get_world_size()
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    return flat.numel()


def min_over_ranks(value, device):
    """ranks own different amounts of experience: every rank must run the same number of all-reduced steps"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return int(t.item())


def broadcast_parameters(model, src=0):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for p in model.state_dict().values():
            dist.broadcast(p, src)


class GraphedStep(object):
    """One optimisation step as CUDA graphs: graph 1 = zero the gradients, forward (control-flow-free ACT,
    value_network.forward_torch(static=True)), MSE, backward; then the flat gradient all-reduce of the data-parallel
    run; graph 2 = the optimizer step (Adam with capturable=True).  ~150 tiny kernels per step are replayed with two
    launches instead of being issued one by one from Python.  The warm-up steps capture needs are undone (parameters
    and optimizer state are restored in place), so a graphed run equals an eager run step for step."""

    def __init__(self, model, optimizer, criterion, inputs, values):
        self.model, self.optimizer = model, optimizer
        self.inputs = inputs.detach().clone()
        self.values = values.detach().clone()
        params = [p for p in model.parameters()]
        saved_p = [p.detach().clone() for p in params]
        saved_s = {id(t): t.detach().clone() for st in optimizer.state.values() for t in st.values()
                   if torch.is_tensor(t)}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                optimizer.zero_grad(set_to_none=True)
                out, _ = model.forward_torch(self.inputs, static=True)
                criterion(out, self.values).backward()
                optimizer.step()
        torch.cuda.current_stream().wait_stream(side)
        self.fwd_bwd = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.fwd_bwd):
            out, _ = model.forward_torch(self.inputs, static=True)
            self.loss = criterion(out, self.values)
            self.loss.backward()
        self.opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.opt):
            optimizer.step()
        with torch.no_grad():                      # undo the warm-up and the capture-time step
            for p, q in zip(params, saved_p):
                p.copy_(q)
            for st in optimizer.state.values():
                for t in st.values():
                    if torch.is_tensor(t):
                        t.copy_(saved_s[id(t)]) if id(t) in saved_s else t.zero_()

    def matches(self, inputs, values):
        return inputs.shape == self.inputs.shape and values.shape == self.values.shape and inputs.is_cuda

    def __call__(self, inputs, values):
        self.inputs.copy_(inputs)
        self.values.copy_(values)
        self.fwd_bwd.replay()
        allreduce_gradients(self.model)
        self.opt.replay()
        return self.loss.detach().clone()


class FusedStep(object):
    """One optimisation step as hand-written kernels (csrc/train.cu): aem_train_grad (forward with the
    reference's batch-global ACT halting + MSE + backward, three launches) -> the flat gradient all-reduce of the
    data-parallel run -> aem_adam_step / aem_sgd_step / aem_rmsprop_step (one launch; the three optimizers of trainer.py:24-34).  The live parameters of the module become views into ONE flat
    float32[113752] vector in blob order, so the kernels, the NCCL all-reduce and the lookahead engine's weight upload all
    work on the same memory without gather / scatter copies.  Adam moments are views into flat vectors too, and are
    shared with the torch optimizer's state so optimizer.state_dict() stays meaningful."""

    def __init__(self, model, optimizer, device):
        from aemcarl_b200 import weights as wts
        from aemcarl_b200.engine import get_engine
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("the fused training step is a CUDA kernel; there is no CPU fallback")
        self.model, self.optimizer = model, optimizer
        self.kind = {optim.Adam: "adam", optim.SGD: "sgd", optim.RMSprop: "rmsprop"}[type(optimizer)]
        self.engine = get_engine(dev.index or 0, role="trainer")
        self.engine.set_act_mode(model.act_fixed, model.act_steps)
        self.dropout_seed = int(torch.initial_seed()) & (2 ** 63 - 1)
        self._drop_p = None
        named = dict(model.named_parameters())
        self.flat = torch.empty(wts.NUM_WEIGHTS, dtype=torch.float32, device=dev)
        self.grad = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.step = 0
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.values = None
        self.steps = torch.zeros(1, dtype=torch.int32, device=dev)
        self._params = []
        with torch.no_grad():
            for key, (off, shape) in wts.weight_offsets().items():
                p = named[key]
                n = p.numel()
                view = self.flat[off:off + n].view(shape)
                view.copy_(p.data)
                p.data = view                                   # the module now lives in the flat vector
                st = optimizer.state.get(p)
                if st:                                          # continue from eager / graphed steps
                    if self.kind == "adam":
                        self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
                        self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                        self.step = max(self.step, int(st["step"]))
                    elif self.kind == "sgd" and st.get("momentum_buffer") is not None:
                        self.exp_avg[off:off + n].copy_(st["momentum_buffer"].reshape(-1))
                        self.step = max(self.step, 1)
                    elif self.kind == "rmsprop":
                        self.exp_avg_sq[off:off + n].copy_(st["square_avg"].reshape(-1))
                        self.step = max(self.step, int(st["step"]))
                self._params.append((p, off, n, shape))
        self.export_state()

    def sync_dropout(self):
        """train() mode = the reference's training regime: nn.TransformerEncoderLayer's dropout (p of the module, 0.1) is live
        in the fused step too (aem_set_train_dropout: counter-based masks, one fresh set per step); eval() mode switches it off"""
        p = float(self.model.encoder_layer.dropout.p) if self.model.training else 0.0
        if p != self._drop_p:
            self.engine.set_train_dropout(p, self.dropout_seed)
            self._drop_p = p

    def export_state(self):
        """make the torch optimizer's state describe the flat vectors (views, no copies).  The step count is ONE shared
        tensor that sync_step() keeps current, and the SGD momentum buffer is always the live view (a zero buffer before the
        first step gives buf = grad, torch's first-step rule), so optimizer.state_dict() and a later fall-back to the torch
        step see the state the kernels left."""
        self._step_t = torch.tensor(float(self.step))
        for p, off, n, shape in self._params:
            m, v = self.exp_avg[off:off + n].view(shape), self.exp_avg_sq[off:off + n].view(shape)
            if self.kind == "adam":
                self.optimizer.state[p] = {"step": self._step_t, "exp_avg": m, "exp_avg_sq": v}
            elif self.kind == "sgd":
                self.optimizer.state[p] = {"momentum_buffer": m}
            else:
                self.optimizer.state[p] = {"step": self._step_t, "square_avg": v}

    def sync_step(self):
        self._step_t.fill_(float(self.step))

    def run_batches(self, memory, batch_size, num_batches, use_graph=False):
        """num_batches iterations of optimize_batch (trainer.py:66-86) on a DeviceReplayMemory with Adam on ONE rank: every
        per-step quantity (draw counter, step count, bias corrections, loss sum) lives on the device, so an iteration is
        one C call (aem_train_step_adam: sample + forward / backward + update) -- or, with use_graph, one replay of the CUDA
        graph that call was captured into (re-captured when the ring size or the learning rate changes)."""
        g = self.optimizer.param_groups[0]
        dev = self.flat.device
        self.sync_dropout()
        k = min(int(batch_size), len(memory))
        n = memory.states.shape[1]
        if getattr(self, "_dev_state", None) is None:
            self._dev_state = torch.zeros(2, dtype=torch.int64, device=dev)
            self._adam_scal = torch.zeros(2, dtype=torch.float32, device=dev)
            self._loss_sum = torch.zeros(1, dtype=torch.float32, device=dev)
        if getattr(self, "_batch_states", None) is None or self._batch_states.shape != (k, n, 13):
            self._batch_states = torch.empty((k, n, 13), dtype=torch.float32, device=dev)
            self._batch_values = torch.empty((k,), dtype=torch.float32, device=dev)
            self._graph_key = None
        self._dev_state[1] = self.step                      # continue from steps taken through __call__ / torch
        self._dev_state[0] = memory._samples_drawn + 1      # DeviceReplayMemory.sample() numbers its draws from 1
        self._loss_sum.zero_()

        def one():
            self.engine.train_step_adam(memory.states, memory.values, len(memory), k, memory.sample_seed, self._dev_state,
                                        self._adam_scal, self.flat, self.grad, self.exp_avg, self.exp_avg_sq, g["lr"], g["betas"],
                                        g["eps"], self._batch_states, self._batch_values, self.loss, self._loss_sum, self.steps)

        done = 0
        if use_graph and num_batches > 2:
            one()                                           # (re)allocates the library's scratch outside any capture
            done += 1
            # the graph holds the ring, the batch buffers AND the engine's scratch pointers: the scratch generation is part
            # of the key, so a larger batch / n that went through this (process-wide) engine in between forces a re-capture
            key = (len(memory), k, n, g["lr"], tuple(g["betas"]), g["eps"], memory.states.data_ptr(), memory.sample_seed,
                   self.engine.train_scratch_generation(), self._drop_p)
            if self._graph_key != key:
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):
                    one()
                done += 1                                   # the capture pass does not execute; replay it once to keep the count
                self._graph.replay()
                self._graph_key = key
            for _ in range(num_batches - done):
                self._graph.replay()
        else:
            for _ in range(num_batches):
                one()
        self.step += num_batches
        self.sync_step()
        memory._samples_drawn += num_batches
        self.model.note_external_update()
        return self._loss_sum[0] / num_batches

    def __call__(self, inputs, values):
        g = self.optimizer.param_groups[0]
        inputs = inputs.detach().to(torch.float32).contiguous()
        targets = values.detach().to(torch.float32).reshape(-1).contiguous()
        if self.values is None or self.values.numel() != inputs.shape[0]:
            self.values = torch.empty(inputs.shape[0], dtype=torch.float32, device=inputs.device)
        self.sync_dropout()
        self.engine.train_grad(self.flat, inputs, targets, self.grad, self.loss, self.values, self.steps)
        world = 1
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            world = dist.get_world_size()
            ev = getattr(self, "allreduce_events", None)          # bench.py --config 4: CUDA events around the NCCL call
            if ev is not None:
                ev.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
                ev[-1][0].record()
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM)
            if ev is not None:
                ev[-1][1].record()
        self.step += 1
        if self.kind == "adam":
            self.engine.adam_step(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, g["lr"], g["betas"], g["eps"], self.step,
                                  1.0 / world)
        elif self.kind == "sgd":
            self.engine.sgd_step(self.flat, self.grad, self.exp_avg, g["lr"], g["momentum"], self.step, 1.0 / world)
        else:
            self.engine.rmsprop_step(self.flat, self.grad, self.exp_avg_sq, g["lr"], g["alpha"], g["eps"], 1.0 / world)
        self.sync_step()
        self.model.note_external_update()
        return self.loss[0].clone()


class Trainer(object):
    def __init__(self, model, memory, device, batch_size, cuda_graph=False, fused=False):
        self.cuda_graph = bool(cuda_graph)
        self.fused = bool(fused)
        self._fused = None
        self._graphed = None
        self.model = model
        self.device = device
        self.criterion = nn.MSELoss().to(device)
        self.memory = memory
        self.data_loader = None
        self.batch_size = batch_size
        self.optimizer = None

    def set_learning_rate(self, learning_rate, optimizer_type="SGD"):
        logging.info("Current learning rate: %f,Optimizer Type %s", learning_rate, optimizer_type)
        self._graphed = None
        self._fused = None
        if optimizer_type == "Adam":
            self.optimizer = optim.Adam(self.model.parameters(), lr=learning_rate, betas=(0.9, 0.99), eps=1e-03,
                                        capturable=self.cuda_graph)
        elif optimizer_type == "RMSprop":
            self.optimizer = optim.RMSprop(self.model.parameters(), lr=learning_rate, alpha=0.9)
        else:
            if optimizer_type != "SGD":
                print("Not support optimizer: ", optimizer_type)
            self.optimizer = optim.SGD(self.model.parameters(), lr=learning_rate, momentum=0.9, weight_decay=0.00)

    def _step(self, inputs, values):
        """one optimisation step; returns the loss as a 0-dim tensor (the callers synchronise once per call, not per step)"""
        if self.fused:
            if getattr(self.model, "act_fixed", False) and int(self.model.act_steps) > 3:
                raise NotImplementedError("the fused training step keeps at most 3 GRU steps (max_ponder of the adaptive "
                                          "network); use the eager / graphed step for act_fixed with more")
            if self._fused is None:
                self._fused = FusedStep(self.model, self.optimizer, inputs.device)
            return self._fused(inputs, values)
        if self.cuda_graph and isinstance(self.optimizer, optim.Adam):
            if self._graphed is None or not self._graphed.matches(inputs, values):
                self._graphed = GraphedStep(self.model, self.optimizer, self.criterion, inputs, values)
            return self._graphed(inputs, values)
        self.optimizer.zero_grad()
        outputs, _ = self.model(inputs)
        loss = self.criterion(outputs, values)
        loss.backward()
        allreduce_gradients(self.model)
        self.optimizer.step()
        return loss.detach()

    def _device_memory(self):
        """DeviceReplayMemory (utils/vec_explorer.py): mini-batches are gathered on the device, no DataLoader"""
        return hasattr(self.memory, "sample") and hasattr(self.memory, "epoch_batches")

    def optimize_epoch(self, num_epochs):
        if self.optimizer is None:
            raise ValueError("Learning rate is not set!")
        fast = self._device_memory()
        if not fast and self.data_loader is None:
            self.data_loader = DataLoader(self.memory, self.batch_size, shuffle=True, drop_last=True)
        average_epoch_loss = 0
        n_batches = len(self.memory) // self.batch_size if fast else len(self.data_loader)
        steps_per_epoch = min_over_ranks(n_batches, self.device)
        for epoch in range(num_epochs):
            epoch_loss = 0
            batches = self.memory.epoch_batches(self.batch_size, drop_last=True) if fast else self.data_loader
            for i, (inputs, values) in enumerate(batches):
                if i >= steps_per_epoch:
                    break
                epoch_loss = epoch_loss + self._step(inputs, values)
            average_epoch_loss = float(epoch_loss) / len(self.memory)
            logging.debug("Average loss in epoch %d: %.2E", epoch, average_epoch_loss)
        return average_epoch_loss

    def optimize_batch(self, num_batches):
        if self.optimizer is None:
            raise ValueError("Learning rate is not set!")
        fast = self._device_memory()
        if fast and self.fused and isinstance(self.optimizer, optim.Adam) and self.memory.device.type == "cuda" \
                and not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) \
                and not (getattr(self.model, "act_fixed", False) and int(self.model.act_steps) > 3):
            if self._fused is None:
                self._fused = FusedStep(self.model, self.optimizer, self.memory.device)
            average_loss = float(self._fused.run_batches(self.memory, self.batch_size, num_batches, use_graph=self.cuda_graph))
            logging.debug("Average loss : %.2E", average_loss)
            return average_loss
        if not fast and self.data_loader is None:
            self.data_loader = DataLoader(self.memory, self.batch_size, shuffle=True)
        losses = 0
        for _ in range(num_batches):
            inputs, values = self.memory.sample(self.batch_size) if fast else next(iter(self.data_loader))
            losses = losses + self._step(inputs, values)
        average_loss = float(losses) / num_batches
        logging.debug("Average loss : %.2E", average_loss)
        return average_loss
