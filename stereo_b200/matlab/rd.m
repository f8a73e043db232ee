% Wrapper function for the roof duality solver -- B200 build.
% Same signature and checks as the reference wrapper (rd.m:3-21); rd_mex (built from rd_mex.cpp
% in this folder) forwards to sb_rd_solve of libstereo_b200.so.
function [solution, energy, lower_bound,num_unlabelled] = rd(U0,U1, E00, E01, E10, E11, connectivity, options)

assert(min(connectivity(:) > 0));
assert( max(connectivity(:)) <= numel(U0) );

% Compile if need be
compile('rd_mex.cpp', 'rd_mex');

% Solve
% Change from matlab from base 1 to base 0.
[solution, energy, lower_bound, num_unlabelled] = rd_mex(U0(:),U1(:), E00(:)', E01(:)', E10(:)', E11(:)', uint32(connectivity-1), options);
