/*
 * JNA mapping of include/mvdecon.h (libmvdecon.so).  Extends the reference's own CUDAFourierConvolution
 * (M/process/cuda/CUDAFourierConvolution.java:26-33) so that an instance can be handed to every place that takes the legacy
 * interface (CUDATools.queryCUDADetails, ComputeBlockSeqThreadCUDAFactory); the additional methods are levels 2 and 3 of
 * INTEGRATION.md.  Not compiled in this repository (no JDK in the build image); the ctypes harness in
 * multiview-reconstruction_b200/__init__.py binds the same symbols with the same signatures and is what the tests run.
 */
package net.preibisch.mvrecon.process.cuda;

import java.util.Arrays;
import java.util.List;

import com.sun.jna.Callback;
import com.sun.jna.Pointer;
import com.sun.jna.Structure;
import com.sun.jna.ptr.IntByReference;
import com.sun.jna.ptr.PointerByReference;

public interface MvDeconB200 extends CUDAFourierConvolution
{
	/** DeconViewPSF.PSFTYPE.ordinal() is the value the library expects (mvdecon.h MVD_PSF_*) */
	public static class Config extends Structure
	{
		public int device;
		public int[] dims = new int[ 3 ];
		public int num_views;
		public int psf_type;
		public float lambda;
		public float min_value;
		public int shard_lo, shard_hi, local_z0, local_nz;
		public int max_fft_len;
		/** Threads.numThreads() of the JVM: reproduces AdjustInput.sumImg of this very run (AdjustInput.java:115-119) */
		public int norm_quirk_threads;
		public int shard_y_lo, shard_y_hi, local_y0, local_ny;
		public int exchange_scheme;

		@Override
		protected List< String > getFieldOrder()
		{
			return Arrays.asList( "device", "dims", "num_views", "psf_type", "lambda", "min_value", "shard_lo", "shard_hi", "local_z0",
					"local_nz", "max_fft_len", "norm_quirk_threads", "shard_y_lo", "shard_y_hi", "local_y0", "local_ny", "exchange_scheme" );
		}
	}

	public static class HaloBox extends Structure
	{
		public Pointer base;
		public long row_floats;
		public int nrows, nplanes, y0, y1, z0, z1, hy_lo, hy_hi, hz_lo, hz_hi;

		@Override
		protected List< String > getFieldOrder()
		{
			return Arrays.asList( "base", "row_floats", "nrows", "nplanes", "y0", "y1", "z0", "z1", "hy_lo", "hy_hi", "hz_lo", "hz_hi" );
		}
	}

	public interface ExchangeFn extends Callback { int invoke( Pointer user, int which, HaloBox box ); }
	public interface ReduceFn extends Callback { int invoke( Pointer user, Pointer values, int count, int op ); }

	String mvd_last_error();
	int mvd_version();
	int mvd_reference_threads();

	// ---- level 2: ComputeBlockSeqThread.runIteration on one halo'd block (mvdecon.h, "L2") ----
	int mvd_block_iteration( int device, float[] psiBlock, float[] imgBlock, float[] weightBlock, int[] blockDimsXYZ,
			float[] kernel1, int[] k1DimsXYZ, float[] kernel2, int[] k2DimsXYZ,
			float lambda, float minValue, float maxIntensity, double[] stats );

	// ---- level 3: resident context ("L3") ----
	int mvd_create( Config cfg, PointerByReference ctx );
	int mvd_destroy( Pointer ctx );
	int mvd_set_view( Pointer ctx, int v, float[] img, float[] weight );
	int mvd_set_psf( Pointer ctx, int v, float[] psf, int[] kdimsXYZ );
	int mvd_set_kernels( Pointer ctx, int v, float[] k1, int[] k1dims, float[] k2, int[] k2dims );
	int mvd_init_views( Pointer ctx );
	int mvd_set_psi( Pointer ctx, float[] psi );
	int mvd_get_psi( Pointer ctx, float[] psi );
	int mvd_set_max_intensities( Pointer ctx, float[] maxPerView );
	int mvd_psi_init( Pointer ctx, int type, double sigma, double[] avgOut, float[] maxOut );
	int mvd_psi_init_from_file( Pointer ctx, String path, int precise, double[] avgOut, float[] maxOut );
	int mvd_make_blending_weights( Pointer ctx, int v, int[] boxMin, int[] boxMax, float[] border, float[] blending );
	int mvd_make_blending_weights_affine( Pointer ctx, int v, int[] imgMin, int[] imgMax, float[] border, float[] blending,
			double[] invAffine, int[] bboxOffset );
	int mvd_normalize_weights( Pointer ctx, double osemSpeedup, int additionalSmooth, float maxDiffRange, float scalingRange );
	int mvd_skip_empty_tiles( Pointer ctx, int on, int[] skippedOut );
	int mvd_run_view_update( Pointer ctx, int v, double[] stats );
	int mvd_run_iterations( Pointer ctx, int n, double[] stats );
	int mvd_run_iteration_mul( Pointer ctx, double[] stats );
	int mvd_tiff_write( String path, float[] data, int[] dimsXYZ );
	int mvd_n5_dims( String datasetDir, int[] dimsXYZ );
	int mvd_n5_read( String datasetDir, float[] out );
	int mvd_n5_write( String datasetDir, float[] data, int[] dimsXYZ, int[] blockSizeXYZ, int gzipLevel );
	int mvd_zarr_write( String path, float[] data, int[] dimsXYZ, int[] chunkSizeXYZ, int gzipLevel, double[] voxelSizeXYZ );

	// ---- multi-GPU: one context per device, halos exchanged by the library ----
	int mvd_comm_unique_id( byte[] id128 );
	int mvd_comm_create( byte[] id128, int world, int rank, int device, PointerByReference comm );
	int mvd_comm_destroy( Pointer comm );
	int mvd_comm_attach( Pointer ctx, Pointer comm, int py, int pz );
	int mvd_exchange_halos( Pointer ctx );
	int mvd_exchange_transport( Pointer ctx, IntByReference transport );
	int mvd_set_exchange_callback( Pointer ctx, ExchangeFn fn, Pointer user );
	int mvd_set_reduce_callback( Pointer ctx, ReduceFn fn, Pointer user );
	int mvd_halo_planes( Pointer ctx, IntByReference lo, IntByReference hi );
	int mvd_halo_rows( Pointer ctx, IntByReference lo, IntByReference hi );
}
