% Compile a gateway only if needed (same role as the reference's compile.m:2-48): the mex file
% is rebuilt when it is missing or older than its source.  The gateways have no solver sources
% of their own -- they link against libstereo_b200.so (built by `make -C stereo_b200/csrc`).
function compile(cpp_file, out_file)
my_path = fileparts(mfilename('fullpath'));
root = fileparts(fileparts(my_path));
lib_dir = fullfile(root, 'stereo_b200', 'lib');
inc_dir = fullfile(root, 'include');

mex_file_name = fullfile(my_path, [out_file '.' mexext]);
cpp_file_name = fullfile(my_path, cpp_file);
mex_file = dir(mex_file_name);
src_file = dir(cpp_file_name);
hdr_file = dir(fullfile(my_path, 'sb_mex_common.h'));

compile_file = isempty(mex_file) || mex_file.datenum < src_file.datenum || mex_file.datenum < hdr_file.datenum;
if compile_file
	if ~exist(fullfile(lib_dir, 'libstereo_b200.so'), 'file')
		error('libstereo_b200.so not found in %s: run make -C stereo_b200/csrc first (there is no CPU fallback).', lib_dir);
	end
	mex(cpp_file_name, '-outdir', my_path, ['-I' inc_dir], ['-I' my_path], ['-L' lib_dir], '-lstereo_b200', ...
		['LDFLAGS=$LDFLAGS -Wl,-rpath,' lib_dir]);
end
