function I = denoiseImage_PnP_ADMM(A, net, onnx_dagnetwork, residual_noise)
% Drop-in for main_files/utils/denoiseImage_PnP_ADMM.m; `net` is a qmri_unetres_load handle.
narginchk(2, 4);
if nargin < 3, onnx_dagnetwork = true; end
if nargin < 4, residual_noise = false; end
validateattributes(A, {'single','double'}, {'nonempty','nonsparse','real','nonnan','finite'}, mfilename, 'A');
if ~islogical(onnx_dagnetwork) && ~isequal(onnx_dagnetwork, 0) && ~isequal(onnx_dagnetwork, 1)
    disp('Error: onnx_dagnetwork not set for denoiseImage_PnP_ADMM()'); return
end
res = qmri_b200_mex('denoise', net, A);
res = reshape(res, size(A, 1), size(A, 2), 10, []);
if residual_noise == true
    I = A - res;
elseif residual_noise == false
    I = res;
else
    disp('Error: residual_noise not set for denoiseImage_PnP_ADMM()'); return
end
end
