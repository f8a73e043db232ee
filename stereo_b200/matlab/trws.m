% Wrapper function for the TRW-S solver -- B200 build.
% Same signature and checks as the reference wrapper (trws.m:2-33); the mex gateway it calls
% (trws_mex, built from trws_mex.cpp in this folder) forwards to sb_trws_solve of
% libstereo_b200.so instead of the CPU solver.
function [solution, energy, lower_bound, iterations] =  ...
			trws(kernel, unary, connectivity, q, qprim, alphas, tol, options)

assert(min(connectivity(:) > 0));
assert( max(connectivity(:) <= numel(unary)))
kernel = int32(kernel);

if (any(isnan(q(:))))
    error('q contains NaN');
end

if (any(isnan(qprim(:))))
    error('qprim contains NaN');
end

% Compile if need be
compile('trws_mex.cpp', 'trws_mex');

% Solve
% Change from matlab from base 1 to base 0.
[solution, energy, lower_bound, iterations] = trws_mex(kernel, unary, uint32(connectivity-1), q, qprim, alphas(:), tol, options);
