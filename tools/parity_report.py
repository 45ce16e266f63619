"""Per-parameter parity report of one training iteration on the GPU (run on the B200 box):

  python tools/parity_report.py [--fixture step_m4_b2_full] [--out gpurun_out/parity.txt]

  * fp32 mode (CUDA-core convolutions) against the reference digests of tests/golden/<fixture>.pt: for every parameter the
    largest error of the strided gradient sample relative to the sample's scale, and the abs-sum error;
  * bf16 mode (the product: tcgen05 convolutions, bf16 activations) against the CPU oracle run here on the same weights /
    inputs / eps / (i, j): relative L2 error and cosine per parameter, losses, images.
The oracle is the checker here (test infrastructure), never the thing measured."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tests.conftest import load_golden
from tests.helpers import golden_state, golden_inputs
import rd_b200.config as rd_config
import rd_b200.kernels as K
import rd_b200.ops as ops
from rd_b200.trainer import Trainer, build_model


def setup(fx, precision):
    cfg = rd_config.default_config(precision=precision)
    cfg.update(fx["cfg"])
    cfg["precision"] = precision
    cfg = rd_config.derive(cfg)
    for k in ("input_output_act", "target_output_act"):
        if k in fx["cfg"]:
            cfg[k] = fx["cfg"][k]
    model = build_model(cfg, "cuda:0")
    model.load_state_dict(golden_state(fx, model))
    model.train(fx["training"])
    tr = Trainer(model, cfg, fx["B"], use_graph=False)
    batch, eps = golden_inputs(fx)
    tr.load_batch(batch, eps, tuple(fx["pair"]))
    return cfg, model, tr, batch, eps


def sample_of(t, n):
    x = t.detach().to("cpu", torch.float64).reshape(-1)
    for n_req in (192, 64, 32):
        st = max(1, x.numel() // n_req)
        s = x[::st][:n_req]
        if s.numel() == n:
            return s.float(), x
    return x[::max(1, x.numel() // n)][:n].float(), x


def run_backward(tr, out):
    out["losses"]["all"].backward()
    ops.flush_mix_bwd()
    fp = tr.fp
    K.grad_norm(fp.grad, fp.segments, fp.nseg, fp.partial, fp.scalars, 1.0)
    gn = float(fp.scalars[0])
    K.grad_scale(fp.grad, fp.segments, fp.nseg, fp.scalars)
    return gn


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fixture", default="step_m4_b2_full")
    ap.add_argument("--out", default="")
    ap.add_argument("--skip-bf16", action="store_true")
    a = ap.parse_args()
    lines = []

    def emit(s):
        print(s)
        lines.append(s)

    fx = load_golden(a.fixture + ".pt")
    # ---------------------------------------------------------------- fp32 vs reference digests
    cfg, model, tr, batch, eps = setup(fx, "fp32")
    out = tr.forward_losses(with_y=fx["with_y"], keep=True)
    gn = run_backward(tr, out)
    emit("== fp32 (CUDA-core parity mode) vs reference digests, fixture %s" % a.fixture)
    for k, v in fx["losses"].items():
        emit("loss %-12s got %.7f ref %.7f rel %.2e" % (k, float(out["losses"][k]), v, abs(float(out["losses"][k]) - v) / max(abs(v), 1e-12)))
    emit("grad_norm got %.6f ref %.6f rel %.2e" % (gn, fx["grad_norm"], abs(gn - fx["grad_norm"]) / fx["grad_norm"]))
    rows = []
    for n, p in model.named_parameters():
        d = fx["grads"][n]
        if d is None:
            continue
        s, x = sample_of(p.grad, d["sample"].numel())
        scale = max(float(d["sample"].abs().max()), 1e-30)
        err = float((s - d["sample"]).abs().max()) / scale
        aerr = abs(float(x.abs().sum()) - d["abssum"]) / max(d["abssum"], 1e-30)
        rows.append((err, aerr, n, p.numel(), scale))
    # Biases of convolutions that feed a BatchNorm / InstanceNorm have an EXACTLY zero gradient: what the reference (and this code) hold
    # there is summation noise (|g| ~ 1e-11 .. 4e-9 against a gradient norm of ~10), not reproducible by any other summation order.
    # They are listed separately: the floor is 1e-7 of the (clipped: unit) gradient norm, the same floor the tests use.
    noise = sorted([r for r in rows if r[4] < 1e-7], reverse=True)
    rows = sorted([r for r in rows if r[4] >= 1e-7], reverse=True)
    emit("fp32 gradients: %d parameters with a gradient above the 1e-7 noise floor; sample error / sample scale: max %.2e, median %.2e; > 1e-3: %d" %
         (len(rows), rows[0][0], rows[len(rows) // 2][0], sum(1 for r in rows if r[0] > 1e-3)))
    for r in rows[:12]:
        emit("   %.2e (abssum rel %.2e) %-70s n=%d scale %.2e" % r)
    emit("fp32: %d parameters whose reference gradient is below the floor (exactly-zero gradients, summation noise on both sides): largest |g| %.2e, "
         "largest absolute difference %.2e" % (len(noise), max([r[4] for r in noise] or [0.0]), max([r[0] * r[4] for r in noise] or [0.0])))
    del model, tr, out
    torch.cuda.empty_cache()
    if a.skip_bf16:
        return finish(a, lines)
    # ---------------------------------------------------------------- bf16 vs oracle
    from oracle.rd_oracle import RDOracle, clone_state, train_iteration
    orc = RDOracle(clone_state(golden_state(fx)), fx["cfg"], training=True, batched_condconv=True)
    o_losses, o_grads, o_gn, o_t = train_iteration(orc, batch, eps, tuple(fx["pair"]), keep=True)
    for compose in ("0", "1"):
        ops.COMPOSE_OUT = compose == "1"
        cfg, model, tr, batch, eps = setup(fx, "bf16")
        out = tr.forward_losses(keep=True)
        gn = run_backward(tr, out)
        emit("== bf16 (product) vs CPU oracle, COMPOSE_OUT=%s" % compose)
        for k, v in o_losses.items():
            emit("loss %-12s got %.6f oracle %.6f rel %.2e" % (k, float(out["losses"][k]), v, abs(float(out["losses"][k]) - v) / max(abs(v), 1e-12)))
        emit("grad_norm got %.5f oracle %.5f rel %.2e" % (gn, o_gn, abs(gn - o_gn) / o_gn))
        B, M = fx["B"], fx["M"]
        T = out["tensors"]
        for i in range(M):
            x = T["x_fake"][i * B:(i + 1) * B].permute(0, 3, 1, 2).float().cpu()
            y = o_t["x_fake"][i].detach()
            emit("x_fake[%d] rel-L2 %.3e   S[%d] max abs %.3e" % (i, float((x - y).norm() / y.norm()), i,
                 float((T["S"][i * B:(i + 1) * B].permute(0, 3, 1, 2).float().cpu() - o_t["si"][i].detach()).abs().max())))
        rows = []
        for n, p in model.named_parameters():
            g = o_grads[n]
            if g is None:
                continue
            x, y = p.grad.float().cpu().reshape(-1).double(), g.reshape(-1).double()
            rel = float((x - y).norm() / (y.norm() + 1e-30))
            cos = float((x * y).sum() / (x.norm() * y.norm() + 1e-30))
            rows.append((rel, cos, n, p.numel(), float(y.norm())))
        gtot = sum(r[4] ** 2 for r in rows) ** 0.5
        nz = [r for r in rows if r[4] < 1e-6 * gtot]          # exactly-zero gradients (see the fp32 section): noise on both sides
        emit("bf16: %d parameters with |g| < 1e-6 of the gradient norm %.3f (exactly-zero gradients) are excluded from the per-parameter check" % (len(nz), gtot))
        rows = [r for r in rows if r[4] >= 1e-6 * gtot]
        big = [r for r in rows if r[3] >= 256]
        big.sort(reverse=True)
        emit("bf16 gradients: %d parameters >= 256 elements: rel-L2 max %.3f median %.3f; cosine min %.4f; rel-L2 > 0.1: %d; cos < 0.995: %d" %
             (len(big), big[0][0], big[len(big) // 2][0], min(r[1] for r in big), sum(1 for r in big if r[0] > 0.1),
              sum(1 for r in big if r[1] < 0.995)))
        for r in big[:12]:
            emit("   rel-L2 %.3f cos %.4f %-70s n=%d |g| %.2e" % r)
        small = sorted([r for r in rows if r[3] < 256], reverse=True)
        emit("   small parameters (< 256 elements): %d, rel-L2 max %.3f" % (len(small), small[0][0] if small else 0.0))
        for r in small[:8]:
            emit("   rel-L2 %.3f cos %.4f %-70s n=%d |g| %.2e" % r)
        del model, tr, out
        torch.cuda.empty_cache()
    ops.COMPOSE_OUT = False
    finish(a, lines)


def finish(a, lines):
    if a.out:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        with open(a.out, "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
