function [S] = setup_subsampling_epi(N, M, percentage, V)
% Drop-in for main_files/subsampling_patterns/setup_subsampling_epi.m (same signature, same returned handles S.for / S.adj).
S.handle = qmri_b200_mex('op_epi', N, M, percentage, double(real(V)));
S.size = [N, M, size(V, 2)];
h = S.handle; n = N * M * size(V, 2);
S.for = @(x) qmri_b200_mex('p_for', h, x);
S.adj = @(x) qmri_b200_mex('p_adj', h, x, n);
end
