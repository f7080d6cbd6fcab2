"""SURVEY 8f row 3 (minimal form): the reference's joint training step (train.py:200-248) runs against the drop-in
classes -- ``lstm`` and ``GPRegressionLayer1`` delegate to torch ops with autograd in train() mode -- and the
eval-mode kernels pick the updated weights up afterwards."""
import pytest
import torch

from oracle import gp_ref, lstm_ref
from util import relerr

pytestmark = pytest.mark.gpu


def test_reference_training_step_runs_on_the_dropin_classes():
    import dvg_b200
    dvg_b200.install_dropin()
    import gpytorch                                  # the real library, or the shim install_dropin() provides
    from dvg_b200.convnets import make_codec
    from models.gp_models import GPRegressionLayer1  # the reference's import path (train.py:13)
    from models.lstm import lstm
    torch.manual_seed(3)
    g_dim, B, n_past, n_future = 90, 8, 2, 3
    encoder, decoder = make_codec("dcgan_64", g_dim, 1)
    encoder, decoder = encoder.cuda(), decoder.cuda()
    frame_predictor = lstm(g_dim, g_dim, 256, 2, B).cuda()
    gp_layer = GPRegressionLayer1(num_dims=g_dim).cuda()
    likelihood = gpytorch.likelihoods.GaussianLikelihood(batch_size=g_dim).cuda()
    opts = [torch.optim.Adam(m.parameters(), lr=0.002) for m in (frame_predictor, encoder, decoder)]
    opts.append(torch.optim.Adam([{"params": gp_layer.parameters()}, {"params": likelihood.parameters()}], lr=0.002))
    mll = gpytorch.mlls.VariationalELBO(likelihood, gp_layer, num_data=B, combine_terms=True)
    mse = torch.nn.MSELoss()
    x = [torch.rand(B, 1, 64, 64, device="cuda") for _ in range(n_past + n_future)]
    losses = []
    for it in range(3):
        for m in (gp_layer, likelihood, frame_predictor, encoder, decoder):
            m.train()
            m.zero_grad()
        with gpytorch.settings.max_cg_iterations(45):
            frame_predictor.hidden = frame_predictor.init_hidden()
            mse_latent = ae_mse = mse_x = mse_gp = max_ll = 0
            for i in range(1, n_past + n_future):                       # train.py:213-234
                h = encoder(x[i - 1])
                h_target = encoder(x[i])[0]
                if i < n_past:
                    h, skip = h
                else:
                    h = h[0]
                h_pred = frame_predictor(h)
                mse_latent = mse_latent + mse(h_pred, h_target)
                gp_pred = gp_layer(h.transpose(0, 1).view(g_dim, B, 1))
                max_ll = max_ll - mll(gp_pred, h_target.transpose(0, 1))
                x_pred = decoder([h_pred, skip])
                ae_mse = ae_mse + mse(decoder([h_target, skip]), x[i])
                mse_gp = mse_gp + mse(decoder([gp_pred.mean.transpose(0, 1), skip]), x[i])
                mse_x = mse_x + mse(x_pred, x[i])
            loss = 1000 * ae_mse + 0.001 * mse_x + 0.01 * mse_latent + 0.001 * mse_gp + 0.0001 * max_ll.sum()   # :239
        loss.backward()
        for m in (frame_predictor, gp_layer, likelihood):
            for n_, p_ in m.named_parameters():
                assert p_.grad is not None and torch.isfinite(p_.grad).all(), n_
        assert gp_layer.variational_strategy.inducing_points.grad.abs().max() > 0
        for o in opts:
            o.step()
        losses.append(loss.item())
    assert all(torch.isfinite(torch.tensor(losses))) and losses[-1] < losses[0], losses
    # eval mode: the kernels must see the weights the optimizers just wrote (version counters -> refresh)
    for m in (gp_layer, likelihood, frame_predictor, encoder, decoder):
        m.eval()
    with torch.no_grad():
        h = encoder(x[0])[0]
        frame_predictor.hidden = frame_predictor.init_hidden()
        y = frame_predictor(h)
        pred = likelihood(gp_layer(h.transpose(0, 1).view(g_dim, B, 1)))
        mean, var = pred.mean, pred.variance
    sd = {k: v.detach().cpu() for k, v in frame_predictor.state_dict().items()}
    y_ref, _ = lstm_ref.lstm_forward(sd, h.cpu(), lstm_ref.init_hidden(2, B, 256))
    assert relerr(y, y_ref) < 1e-4
    gsd = {k: v.detach().cpu() for k, v in gp_layer.state_dict().items()}
    lsd = {k: v.detach().cpu() for k, v in likelihood.state_dict().items()}
    ref = gp_ref.predictive(gsd, lsd, gp_ref.latent_to_gp_input(h.cpu()), torch.float64, "direct", full_cov=False)
    assert relerr(var, ref["variance"]) < 1e-4
    assert relerr(mean, ref["mean"]) < 1e-3
