/*
 * Level-2 glue: a ComputeBlockSeqThread whose runIteration is ONE call into libmvdecon.so (mvd_block_iteration): both
 * convolutions, the quotient and the update run as fused sm_100a passes on the device; the block never returns to the host in
 * between.  Same role and constructor shape as ComputeBlockSeqThreadCUDA (ComputeBlockSeqThreadCUDA.java:44-75), whose
 * runIteration (:77-169) round-trips the block through host memory four times.
 */
package net.preibisch.mvrecon.process.deconvolution.iteration.sequential;

import net.imglib2.Cursor;
import net.imglib2.RandomAccessibleInterval;
import net.imglib2.img.array.ArrayImg;
import net.imglib2.img.basictypeaccess.array.FloatArray;
import net.imglib2.type.numeric.real.FloatType;
import net.imglib2.view.Views;
import net.preibisch.mvrecon.process.cuda.Block;
import net.preibisch.mvrecon.process.cuda.CUDADevice;
import net.preibisch.mvrecon.process.cuda.MvDeconB200;
import net.preibisch.mvrecon.process.deconvolution.DeconView;

public class ComputeBlockSeqThreadB200 extends ComputeBlockSeqThreadAbstract
{
	final MvDeconB200 lib;
	final CUDADevice device;
	final float lambda;
	final float[] img, weight;

	public ComputeBlockSeqThreadB200( final float minValue, final float lambda, final int id, final int[] blockSize,
			final MvDeconB200 lib, final CUDADevice device )
	{
		super( minValue, blockSize, id ); // ArrayImg psiBlockTmp (ComputeBlockThreadAbstract.java:47-60)

		this.lib = lib;
		this.device = device;
		this.lambda = lambda;

		final int n = blockSize[ 0 ] * blockSize[ 1 ] * blockSize[ 2 ];
		this.img = new float[ n ];
		this.weight = new float[ n ];
	}

	@Override
	public IterationStatistics runIteration(
			final DeconView view,
			final Block block,
			final RandomAccessibleInterval< FloatType > imgBlock,
			final RandomAccessibleInterval< FloatType > weightBlock,
			final float maxIntensityView,
			final ArrayImg< FloatType, ? > kernel1,
			final ArrayImg< FloatType, ? > kernel2 )
	{
		// the (virtual, zero-extended) cut-outs become plain arrays, x fastest
		materialise( imgBlock, img );
		materialise( weightBlock, weight );

		final float[] psi = array( (ArrayImg< FloatType, ? >)getPsiBlockTmp() );
		final double[] stats = new double[ 2 ];

		final int rc = lib.mvd_block_iteration( device.getDeviceId(), psi, img, weight, getBlockSize(),
				array( kernel1 ), dims( kernel1 ), array( kernel2 ), dims( kernel2 ),
				lambda, getMinValue(), maxIntensityView, stats );

		if ( rc != 0 )
			throw new RuntimeException( "mvd_block_iteration failed: " + lib.mvd_last_error() );

		final IterationStatistics is = new IterationStatistics();
		is.sumChange = stats[ 0 ];
		is.maxChange = stats[ 1 ];
		return is;
	}

	static float[] array( final ArrayImg< FloatType, ? > img )
	{
		return ( (FloatArray)img.update( null ) ).getCurrentStorageArray();
	}

	static int[] dims( final ArrayImg< FloatType, ? > img )
	{
		return new int[]{ (int)img.dimension( 0 ), (int)img.dimension( 1 ), (int)img.dimension( 2 ) };
	}

	static void materialise( final RandomAccessibleInterval< FloatType > src, final float[] dst )
	{
		final Cursor< FloatType > c = Views.flatIterable( src ).cursor();
		int i = 0;
		while ( c.hasNext() )
			dst[ i++ ] = c.next().get();
	}
}
