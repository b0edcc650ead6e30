/*
 * Level-3 glue: the resident driver.  Everything MultiViewDeconvolutionSeq.runNextIteration does per iteration
 * (MultiViewDeconvolutionSeq.java:58-180: block copy-in, ComputeBlockSeqThread.runIteration, delayed write-back, statistics)
 * happens inside mvd_run_iterations on data that stays in HBM; the Java side hands the views over once and reads psi back at the end.
 *
 * The constructor runs the reference's PsiInit on the host exactly as today (MultiViewDeconvolution.java:115-135) through super(...).
 * The block factory passed to super is only asked for numParallelBlocks(); no block threads compute anything on this path.
 */
package net.preibisch.mvrecon.process.deconvolution;

import java.util.Date;

import com.sun.jna.Pointer;
import com.sun.jna.ptr.PointerByReference;

import net.imglib2.Cursor;
import net.imglib2.RandomAccessibleInterval;
import net.imglib2.img.ImgFactory;
import net.imglib2.img.array.ArrayImg;
import net.imglib2.img.basictypeaccess.array.FloatArray;
import net.imglib2.type.numeric.real.FloatType;
import net.imglib2.view.Views;
import net.preibisch.mvrecon.Threads;
import net.preibisch.mvrecon.process.cuda.MvDeconB200;
import net.preibisch.mvrecon.process.deconvolution.DeconViewPSF.PSFTYPE;
import net.preibisch.mvrecon.process.deconvolution.init.PsiInitFactory;
import net.preibisch.mvrecon.process.deconvolution.iteration.ComputeBlockThreadFactory;
import net.preibisch.mvrecon.process.deconvolution.iteration.sequential.ComputeBlockSeqThread;
import net.preibisch.legacy.io.IOFunctions;

public class MultiViewDeconvolutionB200 extends MultiViewDeconvolution< ComputeBlockSeqThread >
{
	final MvDeconB200 lib;
	final Pointer ctx;
	final float[] psiArray;
	final double[] stats;

	/**
	 * @param rawPSFs - the PSFs as handed to the DeconViewPSF constructors, i.e. BEFORE DeconViewPSF.init normalised them
	 *                  (the library derives kernel1 / kernel2 itself, DeconViewPSF.java:119-254, in list order like DeconViews.java:62-70)
	 */
	public MultiViewDeconvolutionB200(
			final DeconViews views,
			final int numIterations,
			final PsiInitFactory psiInitFactory,
			final ComputeBlockThreadFactory< ComputeBlockSeqThread > computeBlockFactory,
			final ImgFactory< FloatType > psiFactory,
			final MvDeconB200 lib,
			final int device,
			final PSFTYPE psfType,
			final float lambda,
			final float[][] rawPSFs,
			final int[][] rawPSFDims )
	{
		super( views, numIterations, psiInitFactory, computeBlockFactory, psiFactory );

		this.lib = lib;

		final int nx = (int)psi.dimension( 0 ), ny = (int)psi.dimension( 1 ), nz = (int)psi.dimension( 2 );
		final int numViews = views.getViews().size();

		final MvDeconB200.Config cfg = new MvDeconB200.Config();
		cfg.device = device;
		cfg.dims[ 0 ] = nx; cfg.dims[ 1 ] = ny; cfg.dims[ 2 ] = nz;
		cfg.num_views = numViews;
		cfg.psf_type = psfType.ordinal();
		cfg.lambda = lambda;
		cfg.min_value = MultiViewDeconvolution.minValue;
		cfg.shard_lo = 0; cfg.shard_hi = nz; cfg.local_z0 = 0; cfg.local_nz = nz;
		cfg.max_fft_len = 0;
		// AdjustInput.sumImg of THIS JVM (AdjustInput.java:115-119 with FusionTools.divideIntoPortions over Threads.numThreads())
		cfg.norm_quirk_threads = Threads.numThreads();

		final PointerByReference ref = new PointerByReference();
		check( lib.mvd_create( cfg, ref ) );
		this.ctx = ref.getValue();

		final float[] tmpImg = new float[ nx * ny * nz ];
		final float[] tmpWeight = new float[ nx * ny * nz ];

		for ( int v = 0; v < numViews; ++v )
		{
			final DeconView view = views.getViews().get( v );
			materialise( view.getImage(), tmpImg );
			materialise( view.getWeight(), tmpWeight );
			check( lib.mvd_set_view( ctx, v, tmpImg, tmpWeight ) );
			check( lib.mvd_set_psf( ctx, v, rawPSFs[ v ], rawPSFDims[ v ] ) );
		}

		check( lib.mvd_init_views( ctx ) );
		// tiles in which a view has no weight at all are not computed (DeconView.filterBlocksForContent, DeconView.java:204-274)
		check( lib.mvd_skip_empty_tiles( ctx, 1, null ) );

		this.psiArray = ( (FloatArray)( (ArrayImg< FloatType, ? >)psi ).update( null ) ).getCurrentStorageArray();
		this.stats = new double[ 2 * numViews ];

		if ( initWasSuccessful() )
		{
			check( lib.mvd_set_max_intensities( ctx, max ) );
			check( lib.mvd_set_psi( ctx, psiArray ) );
		}
	}

	@Override
	public void runNextIteration()
	{
		++it;

		IOFunctions.println( "iteration: " + it + " (" + new Date( System.currentTimeMillis() ) + ")" );

		check( lib.mvd_run_iterations( ctx, 1, stats ) );

		for ( int v = 0; v < stats.length / 2; ++v )
			IOFunctions.println( "iteration: " + it + ", view: " + v + " --- sum change: " + stats[ 2 * v ] + " --- max change per pixel: " + stats[ 2 * v + 1 ] );

		// the debug view of MultiViewDeconvolution.runIterations (:153-191) reads psi: keep the host copy current when it is on
		if ( debug && ( it - 1 ) % debugInterval == 0 )
			check( lib.mvd_get_psi( ctx, psiArray ) );
	}

	@Override
	public net.imglib2.img.Img< FloatType > getPSI()
	{
		check( lib.mvd_get_psi( ctx, psiArray ) );
		return psi;
	}

	public void close() { lib.mvd_destroy( ctx ); }

	void check( final int rc )
	{
		if ( rc != 0 )
			throw new RuntimeException( "libmvdecon: " + lib.mvd_last_error() );
	}

	static void materialise( final RandomAccessibleInterval< FloatType > src, final float[] dst )
	{
		final Cursor< FloatType > c = Views.flatIterable( src ).cursor();
		int i = 0;
		while ( c.hasNext() )
			dst[ i++ ] = c.next().get();
	}
}
