function [S] = setup_subsampling_epi(N, M, percentage, V)
% Drop-in for main_files/subsampling_patterns/setup_subsampling_epi.m (same signature).
S.handle = qmri_b200_mex('op_epi', N, M, percentage, double(real(V)));
S.size = [N, M, size(V, 2)];
end
