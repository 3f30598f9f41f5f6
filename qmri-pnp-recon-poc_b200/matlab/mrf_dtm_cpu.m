function out = mrf_dtm_cpu(dict, data, par)
% Drop-in for main_files/dictionary_matching/mrf_dtm_cpu.m (same signature; runs on the GPU).
persistent cached_D cached_handle
if isempty(cached_handle) || ~isequal(size(cached_D), size(dict.D)) || ~isequal(cached_D, dict.D)
    cached_handle = qmri_b200_mex('dict_load', single(real(dict.D)), single(dict.normD(:)), single(dict.lut));
    cached_D = dict.D;
end
datadims = size(data.X);
T = datadims(end);
N = prod(datadims(1:end-1));
x = single(reshape(data.X, [N, T]));                     % mask forced all-true (mrf_dtm_cpu.m:51)
Q = size(dict.lut, 2);
if par.f.verbose; fprintf('Matching data \n'); end
[qmap, pd, mt, dm] = qmri_b200_mex('match', cached_handle, x, Q);
if par.f.Xout
    out.Xfit = reshape(bsxfun(@times, pd .* single(dict.normD(dm)), single(dict.D(dm, :))), [datadims(1:end-1), T]);
    out.X = data.X;
end
if par.f.qout
    out.qmap = reshape(qmap, [datadims(1:end-1), Q]);
    out.mask = true([datadims(1:end-1), 1]);
end
if par.f.pdout, out.pd = reshape(pd, [datadims(1:end-1), 1]); end
if par.f.mtout, out.mt = reshape(mt, [datadims(1:end-1), 1]); end
if par.f.dmout, out.dm = reshape(single(dm), [datadims(1:end-1), 1]); end
if par.f.Yout && isfield(data, 'Y'), out.Y = data.Y; end
end
