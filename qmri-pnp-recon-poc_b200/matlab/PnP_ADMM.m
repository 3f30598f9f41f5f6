function [x] = PnP_ADMM(y, param)
% Drop-in for main_files/algorithms/PnP_ADMM/PnP_ADMM.m (same signature and param fields:
% iter, gamma, cg_tol, F, gt_tsmi, X0, net, denoiser_type, noise_map).
% param.net is either a handle returned by qmri_unetres_load (whole loop on the GPU) or any function handle @(v)
% (pluggable prox, host hop).  param.F: the struct from qmri_fft_operator (carries the operator handle), or the reference's
% own closures over P.for / P.adj (main_recon_tsmis_FFT.m:228-229) - then the operator built last by setup_subsampling_* is
% used, after checking on a probe image that param.F.forward really is that operator.
if ~isfield(param, 'denoiser_type'), param.denoiser_type = 'single_level'; end
if isfield(param.F, 'handle')
    h = param.F.handle;
else
    h = qmri_b200_mex('last_op');
    if h == 0, error('qmri:param', 'param.F carries no operator handle and no operator has been built'); end
    probe = randn(size(param.X0));
    ya = param.F.forward(probe); yb = qmri_b200_mex('forward', h, probe);
    if numel(ya) ~= numel(yb) || norm(ya(:) - yb(:)) > 1e-4 * norm(yb(:))
        error('qmri:param', 'param.F is not the operator built by the last setup_subsampling_* call; build F with qmri_fft_operator(P)');
    end
end
x = qmri_b200_mex('pnp_admm', h, double(y), param);
end
