function F = qmri_fft_operator(P)
% Replaces the two closure lines main_recon_tsmis_FFT.m:228-229:
%   F.forward = @(x) P.for(reshape(fft2(x),[],1))/sqrt(N*M);
%   F.adjoint = @(x) (ifft2(reshape(P.adj(x),N,M,[]))*sqrt(N*M));
F.handle = P.handle;
F.forward = @(x) qmri_b200_mex('forward', P.handle, x);
F.adjoint = @(y) qmri_b200_mex('adjoint', P.handle, y, P.size);
end
