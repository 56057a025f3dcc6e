!******************************************************************************************************
! ljmd_c_api -- ISO_C_BINDING interfaces to libljmd.so (include/ljmd.h).
!
! NOTE: this image has no Fortran compiler (no gfortran/flang/f951), so these three files are
! source-only here; they are deliberately tiny and mechanical.  The same C ABI is exercised from
! C++ (examples/md_driver.cpp) and Python (ctypes) in the test-suite.
!******************************************************************************************************
module ljmd_c_api
  use, intrinsic :: iso_c_binding
  implicit none

  ! Mirror of type(sim_params) (scripts/base/md_types.f90:27-50) with guaranteed C layout.
  type, bind(C) :: ljmd_params
    integer(c_int32_t) :: n
    integer(c_int32_t) :: num_cells
    real(c_double)     :: box_length, inv_box_length, volume, density
    real(c_double)     :: dt, dt_half, dt_square_half
    real(c_double)     :: rc, rc_square
    real(c_double)     :: length_scale, inv_length_scale
  end type ljmd_params

  interface
    ! int ljmd_compute_lj(const ljmd_params*, const double* rx, ry, rz, double* ax, ay, az,
    !                     double* epot, double* d_epot, double* dd_epot);
    function ljmd_compute_lj(p, rx, ry, rz, ax, ay, az, epot, d_epot, dd_epot) &
        bind(C, name='ljmd_compute_lj') result(ierr)
      import :: c_int, c_double, ljmd_params
      type(ljmd_params), intent(in) :: p
      real(c_double), intent(in)    :: rx(*), ry(*), rz(*)
      real(c_double), intent(inout) :: ax(*), ay(*), az(*)
      real(c_double), intent(out)   :: epot, d_epot, dd_epot
      integer(c_int) :: ierr
    end function ljmd_compute_lj

    ! int ljmd_verlet_step(const ljmd_params*, double* rx..az, double* epot, ekin, d_epot, dd_epot);
    function ljmd_verlet_step(p, rx, ry, rz, vx, vy, vz, ax, ay, az, epot, ekin, d_epot, dd_epot) &
        bind(C, name='ljmd_verlet_step') result(ierr)
      import :: c_int, c_double, ljmd_params
      type(ljmd_params), intent(in) :: p
      real(c_double), intent(inout) :: rx(*), ry(*), rz(*), vx(*), vy(*), vz(*), ax(*), ay(*), az(*)
      real(c_double), intent(out)   :: epot, ekin, d_epot, dd_epot
      integer(c_int) :: ierr
    end function ljmd_verlet_step

    ! const char* ljmd_last_error(void);
    function ljmd_last_error() bind(C, name='ljmd_last_error') result(msg)
      import :: c_ptr
      type(c_ptr) :: msg
    end function ljmd_last_error
  end interface

contains

  ! Copy sim_params field by field (the Fortran type is not bind(C); its layout is only de facto).
  pure function to_c_params(n, num_cells, box_length, inv_box_length, volume, density, dt, dt_half, &
                            dt_square_half, rc, rc_square, length_scale, inv_length_scale) result(c)
    integer, intent(in) :: n, num_cells
    real(c_double), intent(in) :: box_length, inv_box_length, volume, density, dt, dt_half, &
                                  dt_square_half, rc, rc_square, length_scale, inv_length_scale
    type(ljmd_params) :: c
    c%n = int(n, c_int32_t); c%num_cells = int(num_cells, c_int32_t)
    c%box_length = box_length; c%inv_box_length = inv_box_length; c%volume = volume; c%density = density
    c%dt = dt; c%dt_half = dt_half; c%dt_square_half = dt_square_half
    c%rc = rc; c%rc_square = rc_square
    c%length_scale = length_scale; c%inv_length_scale = inv_length_scale
  end function to_c_params

  ! Reproduce the reference's `stop '<routine>(): <message>'` with the library's message.
  subroutine ljmd_stop_on_error(ierr)
    integer(c_int), intent(in) :: ierr
    character(kind=c_char), pointer :: s(:)
    character(len=512) :: msg
    integer :: i
    if (ierr == 0) return
    call c_f_pointer(ljmd_last_error(), s, [512])
    msg = ''
    do i = 1, 512
      if (s(i) == c_null_char) exit
      msg(i:i) = s(i)
    end do
    write(*,'(a)') trim(msg)
    stop 1
  end subroutine ljmd_stop_on_error

end module ljmd_c_api
