"""CUDA-graph replay of forward + loss + backward (mc_nerf_b200/graph.py) against the eager step.
With the random draws pinned (the same device tensors returned on every call) graph replay must reproduce the
eager loss and gradients; with torch's generator it must draw fresh numbers on every replay."""
import pytest
import torch

from mc_nerf_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
STAGE = "GLOBAL_OPTIM_EPOCH"


class FixedRNG:
    """torch.randn / randperm / Tensor.uniform_ return the same device tensors on every call (cyclic queues)."""

    def __init__(self, rng):
        self.seq = dict(randn=[rng["noise_c"], rng["noise_sel"], rng["noise_f"]], randperm=[rng["perm"]],
                        uniform=[rng["jitter"]])
        self.seq = {k: [t.to(DEV) for t in v] for k, v in self.seq.items()}
        self.pos = dict(randn=0, randperm=0, uniform=0)

    def _next(self, kind):
        t = self.seq[kind][self.pos[kind] % len(self.seq[kind])]
        self.pos[kind] += 1
        return t

    def __enter__(self):
        self._saved = (torch.randn, torch.randperm, torch.Tensor.uniform_)
        me = self
        torch.randn = lambda *size, **kw: me._next("randn").clone()
        torch.randperm = lambda n, **kw: me._next("randperm").clone()
        torch.Tensor.uniform_ = lambda self_t, a=0.0, b=1.0: self_t.copy_(me._next("uniform"))
        return self

    def __exit__(self, *a):
        torch.randn, torch.randperm, torch.Tensor.uniform_ = self._saved


def build(seed=0):
    from mc_nerf_b200.model import MC_Model, MC_NeRF_Loss
    sp = syn.make_sys_param(n_cam=6, img_h=24, img_w=32, batch=256, samples=16, scale=2, device=DEV, with_images=False)
    sp["mlp_precision"] = "bf16"
    torch.manual_seed(seed)
    m = MC_Model(sp).to(DEV)
    with torch.no_grad():
        for k, v in syn.init_camera_weights(sp).items():
            getattr(m, k).copy_(v)
    return sp, m, MC_NeRF_Loss(sp)


def grads(m):
    return {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}


def test_graph_replay_matches_eager_step():
    from mc_nerf_b200.graph import GraphedTrainStep
    # two identically initialised models: the graphed one must never have run an eager backward (an eager autograd
    # graph pins its AccumulateGrad nodes to the default stream, see graph.py)
    sp, m, loss_fn = build()
    _, m_e, loss_fn_e = build()
    rng = syn.draw_step_rng(sp, 256, seed=5)
    batch = tuple(t.to(DEV) for t in syn.make_train_batch(sp, img_id=2, seed=3))
    with FixedRNG(rng):
        step = GraphedTrainStep(m, loss_fn)
        loss_g = step(batch, 25, STAGE, 0.5)
        g_g = grads(m)
        loss_t = loss_fn_e(m_e(batch, 25, STAGE, 0.5)[0], STAGE)
        loss_t.backward()
        loss_e, g_e = loss_t.detach().clone(), grads(m_e)
        assert abs(loss_g.item() - loss_e.item()) <= 1e-6 * abs(loss_e.item())
        assert set(g_g) == set(g_e)
        for k in g_e:     # ray-gradient atomics reorder: tiny differences in the camera-parameter gradients only
            assert torch.allclose(g_g[k], g_e[k], rtol=1e-4, atol=1e-7), k
        # inputs are re-read on replay: another camera / image gives another loss, the first one comes back
        other = tuple(t.to(DEV) for t in syn.make_train_batch(sp, img_id=4, seed=9))
        loss_o = step(other, 25, STAGE, 0.5).item()
        assert abs(loss_o - loss_e.item()) > 1e-6
        assert abs(step(batch, 25, STAGE, 0.5).item() - loss_e.item()) <= 1e-6 * abs(loss_e.item())


def test_graph_replay_follows_weight_updates():
    """The packed bf16 weight images are derived INSIDE the graph: a replay after the fp32 parameters changed (what
    optimizer.step() does) must render with the new weights, exactly as an eager step does."""
    from mc_nerf_b200.graph import GraphedTrainStep
    from mc_nerf_b200.model import RAdam
    sp, m, loss_fn = build()
    _, m_e, loss_fn_e = build()
    rng = syn.draw_step_rng(sp, 256, seed=5)
    batch = tuple(t.to(DEV) for t in syn.make_train_batch(sp, img_id=2, seed=3))

    def perturb(model):
        with torch.no_grad():
            for name, p in model.nerf.named_parameters():      # a change the renders cannot miss
                p.mul_(1.5)
                if name.endswith("bias"):
                    p.add_(0.25)

    with FixedRNG(rng):
        step = GraphedTrainStep(m, loss_fn)
        before = step(batch, 25, STAGE, 0.5).item()
        perturb(m)
        after = step(batch, 25, STAGE, 0.5).item()
        opt = RAdam(list(m.nerf.parameters()), lr=5e-3)      # MLP weights only: the normalised
            # reprojection term makes the camera parameters' first step large and sensitive to summation order
        opt.step()                                  # consumes the replay's gradients, moves every parameter
        after_opt = step(batch, 25, STAGE, 0.5).item()
        perturb(m_e)
        loss = loss_fn_e(m_e(batch, 25, STAGE, 0.5)[0], STAGE)
        loss.backward()
        opt_e = RAdam(list(m_e.nerf.parameters()), lr=5e-3)
        opt_e.step()
        after_opt_e = loss_fn_e(m_e(batch, 25, STAGE, 0.5)[0], STAGE).item()
    assert abs(after - before) > 1e-3 * abs(before), (before, after)        # the replay saw the new weights
    assert abs(after - loss.item()) <= 1e-5 * abs(after)
    assert abs(after_opt_e - loss.item()) > 1e-5 * abs(after_opt_e)         # the eager render re-packed its weights too
    assert abs(after_opt - after_opt_e) <= 1e-4 * abs(after_opt_e), (after, after_opt, after_opt_e)
    assert abs(after_opt - after) > 10 * abs(after_opt - after_opt_e)        # the optimiser's step is what moved it


def test_graph_replay_draws_fresh_random_numbers():
    from mc_nerf_b200.graph import GraphedTrainStep
    sp, m, loss_fn = build(1)
    batch = tuple(t.to(DEV) for t in syn.make_train_batch(sp, img_id=1, seed=3))
    step = GraphedTrainStep(m, loss_fn)
    losses = {round(step(batch, 25, STAGE, 0.5).item(), 7) for _ in range(4)}
    assert len(losses) > 1          # new pixels / jitter / noise on every replay


def test_graph_captures_the_sample_cap_path():
    """samples*scale > 128: the reference's 128-per-ray cap is drawn on the device, so the step still captures and
    every replay draws a fresh subset"""
    from mc_nerf_b200.graph import GraphedTrainStep
    from mc_nerf_b200.model import MC_Model, MC_NeRF_Loss
    sp = syn.make_sys_param(n_cam=4, img_h=16, img_w=16, batch=64, samples=64, scale=4, device=DEV, with_images=False)
    m = MC_Model(sp).to(DEV)
    with torch.no_grad():
        for k, v in syn.init_camera_weights(sp).items():
            getattr(m, k).copy_(v)
    batch = tuple(t.to(DEV) for t in syn.make_train_batch(sp, img_id=1, seed=3))
    step = GraphedTrainStep(m, MC_NeRF_Loss(sp))
    losses = [step(batch, 25, STAGE, 0.5).item() for _ in range(3)]
    assert all(l == l for l in losses) and len(set(losses)) > 1


def test_prefetched_inputs_give_the_same_step():
    from mc_nerf_b200.graph import GraphedTrainStep
    sp, m, loss_fn = build(2)
    rng = syn.draw_step_rng(sp, 256, seed=5)
    host_a = tuple(t.pin_memory() for t in syn.make_train_batch(sp, img_id=2, seed=3))
    host_b = tuple(t.pin_memory() for t in syn.make_train_batch(sp, img_id=4, seed=9))
    with FixedRNG(rng):
        step = GraphedTrainStep(m, loss_fn)
        la = step(host_a, 25, STAGE, 0.5).item()          # direct upload
        lb = step(host_b, 25, STAGE, 0.5).item()
        assert la != lb
        step.prefetch(host_b)
        assert step(host_b, 25, STAGE, 0.5).item() == lb   # staged upload, same inputs -> same loss
        step.prefetch(host_a)
        assert step(host_a, 25, STAGE, 0.5).item() == la
        step.prefetch(host_a)                              # staged A, asked for B: the staging is ignored
        assert step(host_b, 25, STAGE, 0.5).item() == lb
        for _ in range(3):                                 # steady state: prefetch right after every call
            assert step(host_a, 25, STAGE, 0.5).item() == la
            step.prefetch(host_a)


def test_one_graph_follows_the_barf_schedule():
    """cur_ratio changes every step (ref: main.py:80); the BARF weights it selects are read from a device buffer, so the
    same captured graph must reproduce the eager step at every ratio - and the epoch number - without a new capture."""
    from mc_nerf_b200.graph import GraphedTrainStep
    sp, m, loss_fn = build(3)
    _, m_e, loss_fn_e = build(3)
    rng = syn.draw_step_rng(sp, 256, seed=5)
    batch = tuple(t.to(DEV) for t in syn.make_train_batch(sp, img_id=2, seed=3))
    ratios = [0.40, 0.45, 0.52, 0.60, 0.75]          # inside the BARF window (20/52 .. 36/52) and past its end
    with FixedRNG(rng):
        step = GraphedTrainStep(m, loss_fn)
        got = [step(batch, 20 + i, STAGE, r).item() for i, r in enumerate(ratios)]
        assert len(step._graphs) == 1
        want = [loss_fn_e(m_e(batch, 20 + i, STAGE, r)[0], STAGE).item() for i, r in enumerate(ratios)]
    assert len(set(round(x, 6) for x in want[:4])) == 4          # the window really moves the loss
    for a, b in zip(got, want):
        assert abs(a - b) <= 2e-6 * abs(b), (got, want)


@pytest.mark.parametrize("graphed", [False, True])
def test_short_training_run_learns(graphed):
    """End-to-end sanity of the whole loop (render -> loss -> backward -> RAdam -> re-packed weights -> render ...):
    200 steps on a flat-coloured target must cut the rgb loss several-fold, eager and graph-replayed alike."""
    from mc_nerf_b200.graph import GraphedTrainStep
    from mc_nerf_b200.model import RAdam
    sp, m, loss_fn = build(7)
    sp["pixel_sampler"] = "device"
    torch.manual_seed(0)
    gt_img, img_id, iw, ip, ew, ep = syn.make_train_batch(sp, img_id=1, seed=3)
    gt_img = torch.tensor([0.8, 0.3, 0.1]).expand_as(gt_img).contiguous()          # one colour everywhere
    batch = tuple(t.to(DEV) for t in (gt_img, img_id, iw, ip, ew, ep))
    opt = RAdam(list(m.nerf.parameters()), lr=2e-3)
    step = GraphedTrainStep(m, loss_fn) if graphed else None
    rgb_losses = []
    for i in range(200):
        if graphed:
            step(batch, 25, STAGE, 0.9)
        else:
            opt.zero_grad()
            loss_dict = m(batch, 25, STAGE, 0.9)[0]
            loss_fn(loss_dict, STAGE).backward()
        opt.step()
        if i % 20 == 0 or i == 199:
            with torch.no_grad():
                ld = m(batch, 25, STAGE, 0.9)[0]
                rgb_losses.append(float(((ld["rgb"][1] - ld["rgb"][2]) ** 2).mean()))
    assert rgb_losses[-1] < 0.25 * rgb_losses[0], rgb_losses
